"""Host-side geometry of the conv kernels for every layer of every BASELINE configuration, without a GPU: vcb_conv_packed_sizes
runs the same conv_geometry() the launch uses (tile mode, stages, shared-memory / TMEM budget), so a shape that would be
rejected at launch ("not enough shared memory", bad N tile, ...) fails here.  Shapes come from a meta-device pass over the
oracle's YOLOv5 graph (n/s/m/l/x) and from the ReID network's layer table."""
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from torch import nn


def _yolo_convs(name, h, w):
    from oracle import yolov5 as Y
    model = Y.build(name, seed=0).to("meta")
    seen = []
    hooks = [m.register_forward_hook(lambda mod, inp, out: seen.append((tuple(inp[0].shape), mod.in_channels, mod.out_channels,
                                                                         mod.kernel_size[0], mod.stride[0], mod.padding[0])))
             for m in model.modules() if isinstance(m, nn.Conv2d)]
    try:
        model(torch.empty(1, 3, h, w, device="meta"))
    except RuntimeError:
        pass                      # the Detect decode mixes a CPU grid into the meta pass; every convolution has run by then
    for hk in hooks:
        hk.remove()
    det = [m for m in model.modules() if isinstance(m, Y.Detect)][0]
    have = {c[1] for c in seen if c[2] == det.no * det.na}
    for i, stride in enumerate((8, 16, 32)):                 # Detect levels the aborted pass did not reach
        if det.m[i].in_channels not in have:
            seen.append(((1, det.m[i].in_channels, h // stride, w // stride), det.m[i].in_channels, det.no * det.na, 1, 1, 0))
    return seen


@pytest.mark.parametrize("name,h,w", [("yolov5s", 640, 640), ("yolov5m", 640, 640), ("yolov5m", 1024, 1024), ("yolov5l", 384, 640),
                                      ("yolov5l", 736, 1280), ("yolov5n", 96, 160), ("yolov5x", 640, 640)])
def test_every_yolo_layer_has_a_launch_geometry(name, h, w):
    from vehicle_counting_b200 import _lib as L, ops
    convs = _yolo_convs(name, h, w)
    assert len(convs) >= 60 and sum(c[2] == 255 for c in convs) == 3
    checked = 0
    for batch in (1, 8, 64):
        for (shape, cin, cout, k, s, p) in convs:
            hh, ww = shape[2], shape[3]
            variants = []
            if cin == 3:                                   # the 6x6/s2 stem runs as a 3x3/s1 conv over the space-to-depth input
                variants.append(dict(h=hh // 2, w=ww // 2, cin=16, cout=cout, k=3, s=1, p=1, act=L.ACT_SILU))
            elif cout == 255:                              # Detect heads: fp32 logits, no activation
                variants.append(dict(h=hh, w=ww, cin=cin, cout=cout, k=1, s=1, p=0, act=L.ACT_NONE, out_dtype=L.F32, cout_pitch=256))
            else:
                variants.append(dict(h=hh, w=ww, cin=cin, cout=cout, k=k, s=s, p=p, act=L.ACT_SILU))
                if k == 3 and s == 1:                      # Bottleneck with shortcut, in place on a slice of the C3 concat buffer
                    variants.append(dict(h=hh, w=ww, cin=cin, cout=cout, k=k, s=s, p=p, act=L.ACT_SILU, res_mode=L.RES_AFTER_ACT,
                                         res_pitch=2 * cout, cout_pitch=2 * cout))
                if k == 1:                                 # C3: cv1 | cv2 as one GEMM
                    variants.append(dict(h=hh, w=ww, cin=cin, cout=2 * cout, k=1, s=1, p=0, act=L.ACT_SILU))
            for v in variants:
                d = ops.make_conv_desc(batch, v["h"], v["w"], v["cin"], v["cout"], v["k"], v["s"], v["p"], act=v["act"],
                                       res_mode=v.get("res_mode", L.RES_NONE), res_pitch=v.get("res_pitch", 0),
                                       out_dtype=v.get("out_dtype", L.F16), cout_pitch=v.get("cout_pitch"))
                wh, bf = ops.conv_packed_sizes(d)          # raises VcbError with the library's message if the geometry is rejected
                assert wh >= v["cout"] * v["cin"] * v["k"] * v["k"] and bf >= v["cout"]
                ho, wo = ops.conv_out_hw(d)
                assert ho == (v["h"] + 2 * v["p"] - v["k"]) // v["s"] + 1 and wo == (v["w"] + 2 * v["p"] - v["k"]) // v["s"] + 1
                checked += 1
    assert checked > 300


def test_every_reid_layer_has_a_launch_geometry():
    from vehicle_counting_b200 import _lib as L, ops
    layers = [(25, 64, 64, 3, 1, 1)] * 4 + [(25, 64, 128, 3, 2, 1), (25, 64, 128, 1, 2, 0)] + [(13, 128, 128, 3, 1, 1)] * 3 + \
             [(13, 128, 256, 3, 2, 1), (13, 128, 256, 1, 2, 0)] + [(7, 256, 256, 3, 1, 1)] * 3 + \
             [(7, 256, 512, 3, 2, 1), (7, 256, 512, 1, 2, 0)] + [(4, 512, 512, 3, 1, 1)] * 3 + [(50, 16, 64, 3, 1, 1)]
    for crops in (1, 7, 64, 2048, 8192):
        for (hw, cin, cout, k, s, p) in layers:
            for kw in (dict(act=L.ACT_RELU), dict(act=L.ACT_RELU, res_mode=L.RES_BEFORE_ACT, res_pitch=cout), dict(act=L.ACT_NONE, out_dtype=L.F32)):
                d = ops.make_conv_desc(crops, hw, hw, cin, cout, k, s, p, **kw)
                ops.conv_packed_sizes(d)


def test_cv2_linear_table_reproduces_cv2_resize():
    """networks/yolo.py cv2_linear_table + the fixed-point passes of csrc/pointwise.cu letterbox_bilinear_kernel, evaluated in NumPy,
    against cv2.resize(INTER_LINEAR) on uint8: bit-identical for down- and up-scaling ratios."""
    import cv2
    import numpy as np
    from vehicle_counting_b200.networks.yolo import cv2_linear_table
    rng = np.random.default_rng(0)
    cases = [(720, 1280, 360, 640), (720, 1280, 414, 736), (1080, 1920, 360, 640), (200, 300, 213, 320), (100, 150, 320, 480),
             (240, 320, 640, 853), (37, 53, 640, 917), (1000, 30, 640, 19), (480, 640, 480, 640)]
    for _ in range(40):
        sh, sw = (int(v) for v in rng.integers(20, 900, 2))
        r = 640 / max(sh, sw)
        cases.append((sh, sw, max(int(round(sh * r)), 1), max(int(round(sw * r)), 1)))
    for sh, sw, dh, dw in cases:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        xt, yt = cv2_linear_table(dw, sw, False).astype(np.int64), cv2_linear_table(dh, sh, True).astype(np.int64)
        S = src.astype(np.int64)
        H = S[:, xt[:, 0]] * xt[None, :, 2, None] + S[:, xt[:, 1]] * xt[None, :, 3, None]
        out = (((yt[:, 2, None, None] * (H[yt[:, 0]] >> 4)) >> 16) + ((yt[:, 3, None, None] * (H[yt[:, 1]] >> 4)) >> 16) + 2) >> 2
        np.testing.assert_array_equal(np.clip(out, 0, 255).astype(np.uint8), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR),
                                      err_msg=str((sh, sw, dh, dw)))


def test_tuned_layer_table_is_well_formed_and_only_applies_to_measured_shapes(monkeypatch):
    """data/tuned_layers.json (tools/autotune_layers.py on a B200): keys 'k,s,cin,cout,res,HxW,round(log2 n)', values
    [block_n, cta_pair, default us, tuned us] with a >= 3 % gain; the engine applies an entry to exactly that shape and falls back to
    the library's own choice (plus the two small built-in tables) elsewhere; VCB_TUNED=0 switches everything off."""
    import json
    from vehicle_counting_b200 import engine as E
    path = os.path.join(ROOT, "vehicle_counting_b200", "data", "tuned_layers.json")
    layers = json.load(open(path))["layers"]
    assert len(layers) > 50
    for key, v in layers.items():
        k, s, cin, cout, res, hw, lg = key.split(",")
        h, w = hw.split("x")
        assert int(k) in (1, 3) and int(s) in (1, 2) and int(res) in (0, 1) and int(h) > 0 and int(w) > 0 and 0 <= int(lg) <= 13
        bn, cp, us0, us1 = v
        assert bn in (0, 64, 128) and cp in (0, 1, 2, 4, 5) and (bn, cp) != (0, 0)
        assert bn == 0 or (int(cout) % bn == 0 and bn < int(cout))
        assert us1 <= 0.97 * us0 + 0.1                       # the table stores rounded microseconds
    monkeypatch.setattr(E, "_TUNED_TABLE", None)
    key = next(iter(layers))
    k, s, cin, cout, res, hw, lg = key.split(",")
    h, w = (int(v) for v in hw.split("x"))
    n = 2 ** int(lg)
    assert E.tuned_choice(int(k), int(s), int(cin), int(cout), bool(int(res)), n * h * w, n, h, w) == tuple(layers[key][:2])
    assert E.tuned_choice(3, 1, 40, 40, False, 1000, 1, 25, 40) == (0, 0)                     # unknown shape: the library decides
    assert E.tuned_choice(1, 1, 192, 192, False, 5000, 3, 33, 50) == (64, 0)                  # built-in table, small M: no pair override
    monkeypatch.setattr(E, "_TUNED_TABLE", None)
    monkeypatch.setenv("VCB_TUNED", "0")
    assert E.tuned_choice(int(k), int(s), int(cin), int(cout), bool(int(res)), n * h * w, n, h, w) == (0, 0)
    monkeypatch.setattr(E, "_TUNED_TABLE", None)
