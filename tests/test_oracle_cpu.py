"""CPU tests: the oracle against the committed golden vectors (generated from the reference itself by
oracle/make_goldens.py) and against independent checks (torchvision NMS, upstream's published parameter
counts)."""
import os

import numpy as np
import pytest
import torch

from oracle import reid as R
from oracle import ref_shim
from oracle import yolov5 as Y

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CKPT_NPZ = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "reid_ckpt.npz")


def _reid_sd():
    if os.path.isfile(CKPT_NPZ):
        return R.load_state_dict(CKPT_NPZ)
    if ref_shim.available():
        return R.load_state_dict(ref_shim.REID_CKPT)
    pytest.skip("ReID checkpoint repack not present (run python -m oracle.make_goldens in the build container)")


def _golden_crops(z):
    crops, off = [], 0
    for h, w in z["crop_shapes"]:
        crops.append(z["crops_flat"][off:off + h * w * 3].reshape(h, w, 3))
        off += h * w * 3
    return crops


@pytest.mark.parametrize("name,count", [("yolov5n", 1867405), ("yolov5s", 7225885), ("yolov5m", 21172173),
                                        ("yolov5l", 46533693), ("yolov5x", 86705005)])
def test_yolo_fused_param_counts_match_upstream(name, count):
    assert Y.fused_param_count(Y.DetectionModel(name)) == count


def test_yolo_conv_flops_match_survey():
    assert abs(Y.conv_flops(Y.DetectionModel("yolov5s"), 640, 640) / 1e9 - 16.434) < 0.01


def test_yolo_oracle_matches_golden():
    z = np.load(os.path.join(GOLD, "yolo_golden.npz"))
    m = Y.build("yolov5n", seed=0, obj_bias=-1.0)
    dets, pred, raw = Y.autoshape_forward(m, list(z["imgs"]), size=96, return_raw=True)
    for i in range(3):
        np.testing.assert_allclose(raw[i].numpy(), z[f"raw{i}"], rtol=0, atol=2e-4)
    assert [d.shape[0] for d in dets] == z["det_counts"].tolist()
    np.testing.assert_allclose(torch.cat(dets, 0).numpy(), z["dets"], rtol=0, atol=2e-3)


def test_greedy_nms_equals_torchvision():
    import torchvision
    rng = np.random.default_rng(0)
    for n in (1, 17, 400):
        ctr = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(5, 80, (n, 2))
        boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
        scores = rng.uniform(0, 1, n).astype(np.float32)
        mine = Y.greedy_nms(boxes, scores, 0.45)
        tv = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.45).numpy()
        np.testing.assert_array_equal(mine, tv)


def test_letterbox_and_scale_coords_roundtrip():
    shape1 = Y.autoshape_shapes([(720, 1280)], 640)
    assert shape1 == [384, 640]
    (nw, nh), top, bottom, left, right = Y.letterbox_params((720, 1280), shape1)
    assert (nw, nh, top, bottom, left, right) == (640, 360, 12, 12, 0, 0)
    c = torch.tensor([[0.0, 12.0, 640.0, 372.0]])
    out = Y.scale_coords(shape1, c, (720, 1280))
    np.testing.assert_allclose(out.numpy(), [[0, 0, 1280, 720]], atol=1e-4)


def test_adapter_contract_shapes():
    m = Y.build("yolov5n", seed=0, obj_bias=-1.0)
    z = np.load(os.path.join(GOLD, "yolo_golden.npz"))
    out = Y.yolo_backbone_detect(m, {"imgs": list(z["imgs"])}, size=96)
    assert len(out) == 2
    for o in out:
        assert o["bboxes"].dtype == np.float64 and o["bboxes"].shape[1] == 4
        assert (np.diff(o["scores"]) <= 0).all()
    empty = Y.yolo_backbone_detect(Y.build("yolov5n", seed=0, obj_bias=-30.0), {"imgs": list(z["imgs"])}, size=96)
    assert all(o["bboxes"].shape == (0,) for o in empty)


# ------------------------------------------------------------------------------------------ ReID
def test_reid_preprocess_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "reid_golden.npz"))
    got = R.preprocess(_golden_crops(z)).numpy()
    np.testing.assert_allclose(got, z["preprocessed"], rtol=0, atol=2e-6)


def test_reid_forward_matches_reference_golden_train_and_eval():
    z = np.load(os.path.join(GOLD, "reid_golden.npz"))
    sd = _reid_sd()
    crops = _golden_crops(z)
    np.testing.assert_allclose(R.extract(sd, crops, "train"), z["feat_train"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(R.extract(sd, crops[:3], "train"), z["feat_train_first3"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(R.extract(sd, crops, "eval"), z["feat_eval"], rtol=0, atol=5e-6)
    # the reference's embedding depends on the composition of the call (train-mode BN): SURVEY §0.4
    assert np.abs(z["feat_train"][:3] - z["feat_train_first3"]).max() > 1e-2


def test_crop_rule_matches_reference():
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ds = ref_shim.reference_deepsort()
    ds.height, ds.width = 240, 320
    rng = np.random.default_rng(1)
    for _ in range(200):
        x1, y1 = rng.uniform(-20, 300), rng.uniform(-20, 220)
        b = np.array([[x1, y1, x1 + rng.uniform(1, 120), y1 + rng.uniform(1, 120)]])
        xywh = ds._xyxy_to_xywh(b)
        assert ds._xywh_to_xyxy(xywh[0]) == R.crop_box(b[0], 320, 240)


def test_v5_blocks_equal_their_v6_expressions():
    """Focus == 3x3 conv over the space-to-depth input in the ingest's channel order (engine.focus_weights_to_s2d); SPP(5, 9, 13)
    == the SPPF cascade of 5x5 max-pools (what vcb_sppf_pool computes)."""
    import torch
    import torch.nn.functional as F
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import focus_weights_to_s2d, infer_model_name, infer_version
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 12, 16, generator=g)
    w = torch.randn(8, 12, 3, 3, generator=g)
    focus = F.conv2d(torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), 1), w, None, 1, 1)
    n, _, h, wd = x.shape
    s2d = x.view(n, 3, h // 2, 2, wd // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 12, h // 2, wd // 2)     # channel (dy*2+dx)*3+c
    mine = F.conv2d(s2d, focus_weights_to_s2d(w)[:, :12], None, 1, 1)
    assert torch.allclose(focus, mine, atol=1e-5)
    y = torch.randn(2, 4, 13, 9, generator=g)
    mp = lambda t, k: F.max_pool2d(t, k, 1, k // 2)
    assert torch.equal(mp(y, 9), mp(mp(y, 5), 5)) and torch.equal(mp(y, 13), mp(mp(mp(y, 5), 5), 5))
    m5 = Y.build("yolov5n", seed=0, version="v5")
    sd = m5.state_dict()
    assert infer_version(sd) == "v5" and infer_model_name(sd) == "yolov5n"
    assert infer_version(Y.build("yolov5n", seed=0).state_dict()) == "v6"
