"""Host pieces of the per-video loop (no GPU): the mirror of the reference's frame source (modules/datasets.py:14-94) on the FFV1
clip of the pipeline golden, and -- where the reference tree is present (build container) -- equality with the reference's own
VideoLoader batches.  The golden CSV itself is checked for shape (it was written by the unmodified reference driver)."""
import os
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _clip(tmp_path):
    from oracle import make_goldens as M
    z = np.load(os.path.join(GOLD, "pipeline_golden.npz"))
    path = M.write_pipeline_inputs(str(tmp_path), z["base"], int(z["T"]), int(z["step"]))
    return M, z, path


def test_video_loader_mirror_yields_reference_batches(tmp_path):
    from vehicle_counting_b200.modules.datasets import VideoLoader
    M, z, path = _clip(tmp_path)
    cfg = types.SimpleNamespace(image_size=[640, 640], keep_ratio=True)
    frames = M.pipeline_clip_frames(z["base"], int(z["T"]), int(z["step"]))
    got = list(VideoLoader(cfg, path))
    assert len(got) == len(frames)
    for t, b in enumerate(got):
        assert b["frames"] == [t + 1]                                  # 1-based ids (datasets.py:53)
        np.testing.assert_array_equal(b["ori_imgs"][0], frames[t])     # FFV1 is lossless through cv2
        np.testing.assert_array_equal(b["imgs"][0], frames[t][:, :, ::-1])   # the detector sees RGB (datasets.py:55)
    two = list(VideoLoader(cfg, path, batch_size=3))
    assert [len(b["imgs"]) for b in two] == [3, 3, 2]
    from oracle import ref_shim
    if ref_shim.available():
        ref_shim.install()
        from modules.datasets import VideoLoader as RefLoader  # type: ignore
        for b, r in zip(got, RefLoader(cfg, path)):
            assert b["frames"] == r["frames"]
            np.testing.assert_array_equal(b["imgs"][0], r["imgs"][0])
            np.testing.assert_array_equal(b["ori_imgs"][0], r["ori_imgs"][0])


def test_pipeline_golden_csv_shape():
    import pandas as pd
    df = pd.read_csv(os.path.join(GOLD, "pipeline_golden.csv"))
    assert list(df.columns) == ["track_id", "frame_id", "box", "color", "label", "direction", "fpoint", "lpoint", "fframe", "lframe"]
    assert len(df) > 50 and df.frame_id.between(1, 8).all() and (df.fframe <= df.frame_id).all() and (df.frame_id <= df.lframe).all()


def test_video_loader_pinned_ring_request_without_a_gpu_falls_back_to_plain_frames(tmp_path, monkeypatch):
    """$VCB_PINNED_FRAMES asks the frame source for page-locked buffers; without a CUDA device there is nothing to pin, and the
    loader must hand out the same frames as always."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (the pinned path is covered by tests/test_pipeline_gpu.py)")
    from vehicle_counting_b200.modules.datasets import VideoLoader
    M, z, path = _clip(tmp_path)
    cfg = types.SimpleNamespace(image_size=[640, 640], keep_ratio=True)
    frames = M.pipeline_clip_frames(z["base"], int(z["T"]), int(z["step"]))
    monkeypatch.setenv("VCB_PINNED_FRAMES", "4")
    got = list(VideoLoader(cfg, path))
    assert len(got) == len(frames)
    for t, b in enumerate(got):
        np.testing.assert_array_equal(b["ori_imgs"][0], frames[t])
        np.testing.assert_array_equal(b["imgs"][0], frames[t][:, :, ::-1])
