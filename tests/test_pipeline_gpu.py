"""J1: the per-video loop of the reference (modules/__init__.py:28-93) on the GPU path, against the CSV the UNMODIFIED
reference driver wrote for the same clip (tests/golden/pipeline_golden.{npz,csv}, oracle/make_goldens.py:make_pipeline: reference
VideoLoader + ImageDetect + Detector + VideoTracker + VideoCounting + CSV writer, detector network = CPU oracle).

Two runs of the mirror `vehicle_counting_b200.modules.CountingPipeline` on the FFV1 `cam_04.avi` synthesised from the golden's base
frame:
  (a) tracker-input-identical: the detector stage returns the oracle's detections, everything after it is
      the GPU path (frame upload, crop rule, ReID CNN with BatchNorm as shipped, association, zone filter, CSV).  Integer CSV
      fields must be IDENTICAL to the reference's.
  (b) full GPU path (detector on the GPU too): fp16 detections differ from the fp32 oracle's by fractions of a pixel, which can
      flip an int() or an NMS decision, so agreement is asserted as a rate and logged.
"""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CKPT_NPZ = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
INT_COLS = ["track_id", "frame_id", "box", "label", "direction", "fframe", "lframe"]


def _setup(tmp_path):
    from oracle import make_goldens as M
    z = np.load(os.path.join(GOLD, "pipeline_golden.npz"))
    clip = M.write_pipeline_inputs(str(tmp_path), z["base"], int(z["T"]), int(z["step"]))
    tracking = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg_")}
    args = types.SimpleNamespace(weight="seeded", input_path=clip, output_path=str(tmp_path / "out"), mapping=None, mapping_dict=None)
    config = types.SimpleNamespace(model_name=M.PIPE_CFG["model"], min_iou=0.45, min_conf=0.25, max_det=300, image_size=[640, 640], keep_ratio=True)
    cam_config = types.SimpleNamespace(zone_path=str(tmp_path), checkpoint=CKPT_NPZ, cam={"cam_04": {"tracking_config": tracking}})
    return M, float(z["obj_bias"]), args, config, cam_config


def _rows(path):
    import pandas as pd
    df = pd.read_csv(path)
    return [tuple(str(r[c]) for c in INT_COLS) for _, r in df.iterrows()]


def _need_inputs():
    if not (os.path.isfile(os.path.join(GOLD, "pipeline_golden.npz")) and os.path.isfile(CKPT_NPZ)):
        pytest.skip("pipeline golden or shipped ReID weights not present")


def test_pipeline_csv_identical_given_reference_detections(lib, tmp_path, monkeypatch):
    _need_inputs()
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import modules as VM
    from vehicle_counting_b200.modules import detect as VD
    M, obj_bias, args, config, cam_config = _setup(tmp_path)
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    model = Y.build(M.PIPE_CFG["model"], seed=0, obj_bias=obj_bias)

    class OracleBackbone(torch.nn.Module):          # the same substitution the golden used (tests may run the oracle)
        def __init__(self):
            super().__init__()
            self.class_names = [f"class{i}" for i in range(80)]
            self._p = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

        def detect(self, batch, device):
            return M.drop_small_boxes(Y.yolo_backbone_detect(model, batch, size=640, conf=0.25, iou=0.45, max_det=300))

    monkeypatch.setattr(VD, "get_model", lambda a, c: OracleBackbone())
    monkeypatch.setenv("VCB_REID_BN", "train")       # BatchNorm as the reference ships it
    VM.CountingPipeline(args, config, cam_config).run()
    got = _rows(os.path.join(args.output_path, "cam_04.csv"))
    want = _rows(os.path.join(GOLD, "pipeline_golden.csv"))
    assert len(want) > 50
    assert got == want, (len(got), len(want), [r for r in want if r not in got][:3])


def test_pipeline_csv_agreement_full_gpu_path(lib, tmp_path, monkeypatch):
    _need_inputs()
    from oracle import yolov5 as Y
    from vehicle_counting_b200 import modules as VM
    from vehicle_counting_b200.modules import detect as VD
    from vehicle_counting_b200.networks.yolo import YoloBackbone
    M, obj_bias, args, config, cam_config = _setup(tmp_path)
    sd = Y.build(M.PIPE_CFG["model"], seed=0, obj_bias=obj_bias).state_dict()

    class Backbone(YoloBackbone):
        def detect(self, batch, device=None):
            return M.drop_small_boxes(super().detect(batch, device))

    monkeypatch.setattr(VD, "get_model", lambda a, c: Backbone(None, c.min_iou, c.min_conf, c.max_det, state_dict=sd))
    monkeypatch.setenv("VCB_REID_BN", "train")
    VM.CountingPipeline(args, config, cam_config).run()
    got = _rows(os.path.join(args.output_path, "cam_04.csv"))
    want = _rows(os.path.join(GOLD, "pipeline_golden.csv"))
    identical = len(set(got) & set(want))

    def key(r):      # (frame, label) -> list of int boxes
        return (r[1], r[3])
    near = 0
    by = {}
    for r in got:
        by.setdefault(key(r), []).append(np.array(json.loads(r[2])))
    for r in want:
        b = np.array(json.loads(r[2]))
        if any(np.abs(b - g).max() <= 2 for g in by.get(key(r), [])):
            near += 1
    rec = {"reference_rows": len(want), "gpu_rows": len(got), "identical_rows (all integer fields)": identical,
           "rows_with_box_within_2px (same frame, label)": near}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pipeline_csv_agreement.json"), "w") as fh:
        json.dump(rec, fh)
    assert len(got) > 0
    assert near >= 0.6 * len(want), rec


def test_video_loader_pinned_ring_yields_the_same_frames_and_uploads_in_place(lib, tmp_path, monkeypatch):
    """$VCB_PINNED_FRAMES: the frame source decodes into a ring of page-locked buffers (modules/datasets.py mirror); the batches are
    the same bytes as without it and the stage upload takes the in-place path (no gather)."""
    _need_inputs()
    from oracle import make_goldens as M
    from vehicle_counting_b200 import hostcopy
    from vehicle_counting_b200.modules.datasets import VideoLoader
    z = np.load(os.path.join(GOLD, "pipeline_golden.npz"))
    clip = M.write_pipeline_inputs(str(tmp_path), z["base"], int(z["T"]), int(z["step"]))
    cfg = types.SimpleNamespace(image_size=[640, 640], keep_ratio=True)
    plain = [(b["imgs"][0].copy(), b["ori_imgs"][0].copy()) for b in VideoLoader(cfg, clip)]
    monkeypatch.setenv("VCB_PINNED_FRAMES", "4")
    stream = torch.cuda.current_stream()
    n = 0
    for b in VideoLoader(cfg, clip, batch_size=2):
        for k in range(len(b["imgs"])):
            np.testing.assert_array_equal(b["imgs"][k], plain[n][0])
            np.testing.assert_array_equal(b["ori_imgs"][k], plain[n][1])
            n += 1
        dev = torch.zeros((len(b["imgs"]),) + b["imgs"][0].shape, dtype=torch.uint8, device="cuda:0")
        assert hostcopy.upload_inplace(dev, b["imgs"], stream) is True
        torch.cuda.synchronize()
        assert torch.equal(dev.cpu(), torch.from_numpy(np.stack(b["imgs"])))
    assert n == len(plain)
