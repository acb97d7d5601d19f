"""BASELINE.json's full-size configurations on the B200, checked through size-independent properties (the CPU oracle would
need minutes at these sizes; small-size oracle parity is in test_engine_gpu.py / test_kernels_gpu.py / test_conv_gpu.py):
  configs[1]/[3]  YOLOv5m 640x640 batches: batch-position independence (bit-exact), NMS invariants on the real outputs
  configs[2]      YOLOv5m 1024x1024 head tensors against the fp32 oracle on one frame + 64 crops/frame ReID invariants
  configs[4]      YOLOv5l on 1280x720 frames at the reference's inference shape 384x640 (AutoShape size=640)
  K1              linearity of the conv kernel on a full layer, and bit-equality of the kernel variants that share a K order."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _iou(a, b):
    x1, y1 = np.maximum(a[0], b[:, 0]), np.maximum(a[1], b[:, 1]); x2, y2 = np.minimum(a[2], b[:, 2]), np.minimum(a[3], b[:, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter)


def _check_nms_invariants(det, cnt, eng, w, h):
    for b in range(det.shape[0]):
        n = int(cnt[b]); d = det[b, :n]
        assert n <= eng.max_det
        assert np.all(np.diff(d[:, 4]) <= 0), "rows sorted by descending score"
        assert np.all(d[:, 4] > eng.conf)
        assert np.all(d[:, 0] >= 0) and np.all(d[:, 1] >= 0) and np.all(d[:, 2] <= w) and np.all(d[:, 3] <= h)
        assert np.all(d[:, 5] == np.floor(d[:, 5])) and np.all(d[:, 5] >= 0) and np.all(d[:, 5] < eng.nc)
        # no kept box is suppressed by an earlier (higher score) box of its class; NMS ran on the unclipped boxes, so only pairs
        # whose boxes the final clip to the frame did not touch can be re-checked from the output
        inside = (d[:, 0] > 0) & (d[:, 1] > 0) & (d[:, 2] < w) & (d[:, 3] < h)
        for i in range(1, n):
            same = (d[:i, 5] == d[i, 5]) & inside[:i]
            if inside[i] and same.any():
                assert _iou(d[i, :4], d[:i][same][:, :4]).max() <= eng.iou + 1e-4


def test_yolov5m_640_batch_position_independence_and_nms_invariants(lib):
    from vehicle_counting_b200.engine import YoloEngine
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    B, S = 16, 640
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    eng = YoloEngine(synth_yolov5_state_dict("yolov5m", seed=0, obj_bias=-3.0), B, S, S, model_name="yolov5m")
    eng.upload(torch.from_numpy(frames).pin_memory()); eng.forward()
    det, cnt = eng.download()
    det, cnt = det.copy(), cnt.copy()
    logits = [l.clone() for l in eng.logits]
    assert cnt.sum() > 0
    _check_nms_invariants(det, cnt, eng, S, S)
    perm = rng.permutation(B)
    eng.upload(torch.from_numpy(frames[perm]).pin_memory()); eng.forward()
    det2, cnt2 = eng.download()
    for li in range(3):       # a frame's head tensor does not depend on where it sits in the batch: bit-exact
        assert torch.equal(eng.logits[li], logits[li][torch.from_numpy(perm).to(DEV)])
    np.testing.assert_array_equal(cnt2, cnt[perm])
    for b in range(B):
        np.testing.assert_array_equal(det2[b, :cnt2[b]], det[perm[b], :cnt[perm[b]]])


def test_yolov5m_1024_heads_match_oracle_one_frame(lib):
    """configs[2]: YOLOv5m at 1024x1024 (P = 64512 predictions): head tensors of one frame against the fp32 oracle."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import YoloEngine
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    model = Y.build("yolov5m", seed=0, obj_bias=-3.0)
    img = np.random.default_rng(2).integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
    _, _, raw_ref = Y.autoshape_forward(model, [img], size=1024, return_raw=True)
    eng = YoloEngine(model.state_dict(), 1, 1024, 1024, model_name="yolov5m")
    eng.upload(torch.from_numpy(img[None]).pin_memory()); eng.forward()
    det, cnt = eng.download()
    assert sum(l.shape[1] * l.shape[2] * 3 for l in eng.logits) == 64512
    for li in range(3):
        got = eng.logits[li].float().cpu()[..., :3 * eng.no].permute(0, 3, 1, 2)
        rel = ((got - raw_ref[li]).norm() / raw_ref[li].norm()).item()
        assert rel < 7e-3, (li, rel)            # same bar as test_engine_gpu.py (fp16 storage through ~80 convolutions)
    _check_nms_invariants(det, cnt, eng, 1024, 1024)


def test_yolov5l_384x640_runs_and_keeps_invariants(lib):
    """configs[4]: 1280x720 frames reach YOLOv5l at 384x640 in the reference (AutoShape size=640); letterboxed on the host."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import YoloEngine
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    rng = np.random.default_rng(4)
    frames = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(2)]
    lb = np.stack([Y.letterbox(f, (384, 640)) for f in frames])
    assert lb.shape == (2, 384, 640, 3)
    eng = YoloEngine(synth_yolov5_state_dict("yolov5l", seed=0, obj_bias=-3.0), 2, 384, 640, model_name="yolov5l")
    eng.upload(torch.from_numpy(lb).pin_memory()); eng.forward()
    det, cnt = eng.download()
    assert sum(l.shape[1] * l.shape[2] * 3 for l in eng.logits) == 15120
    assert all(torch.isfinite(l).all() for l in eng.logits)
    _check_nms_invariants(det, cnt, eng, 640, 384)


def test_reid_2048_crops_invariants(lib):
    """64 crops/frame x 32 frames through the folded-BN ReID path: unit norm, order independence (bit-exact), duplicates agree."""
    from vehicle_counting_b200.engine import ReidEngine
    from vehicle_counting_b200.weights import synth_reid_state_dict
    rng = np.random.default_rng(8)
    F_, S, n = 32, 640, 2048
    frames = torch.from_numpy(rng.integers(0, 256, (F_, S, S, 3), dtype=np.uint8)).to(DEV)
    wh = rng.uniform(32, 256, (n, 2)); tl = rng.uniform(0, 1, (n, 2)) * (S - wh)
    rois = np.concatenate([np.repeat(np.arange(F_), n // F_)[:, None], tl, tl + wh], 1).astype(np.int32)
    rois[1] = rois[0]                                             # a duplicate crop
    eng = ReidEngine(synth_reid_state_dict(0), capacity=n, bn_mode="eval")
    eng.run(frames, rois)
    f1 = eng.download(n).copy()
    assert np.isfinite(f1).all()
    np.testing.assert_allclose(np.linalg.norm(f1, axis=1), 1.0, atol=1e-3)
    np.testing.assert_array_equal(f1[0], f1[1])
    perm = rng.permutation(n)
    eng.run(frames, rois[perm])
    f2 = eng.download(n)
    np.testing.assert_array_equal(f2, f1[perm])


def test_conv_full_layer_linearity_and_variant_equality(lib):
    """3x3 96->96 over 32 x 80 x 80 (M = 204800, the P3 bottleneck layer of configs[3] at B=32), no activation, fp32 out:
    conv(x1 + x2) == conv(x1) + conv(x2) - bias to fp32 rounding, and the 128-row kernel, the CTA-pair kernel and the
    256-row kernel (same K order, fp32 accumulation in TMEM) agree bit for bit."""
    from vehicle_counting_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(9)
    n, h, w, c = 32, 80, 80, 96
    # values on a coarse grid keep x1 + x2 exactly representable in fp16
    x1 = (torch.randint(-8, 9, (n, h, w, c), generator=g).float() / 8).half().to(DEV)
    x2 = (torch.randint(-8, 9, (n, h, w, c), generator=g).float() / 8).half().to(DEV)
    wt = (torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).half().float().to(DEV)
    bias = (torch.randn(c, generator=g) * 0.5).to(DEV)

    def run(x, cta_pair):
        d = ops.make_conv_desc(n, h, w, c, c, 3, 1, 1, act=L.ACT_NONE, out_dtype=L.F32, cta_pair=cta_pair)
        wp, bp = ops.pack_conv_weights(d, wt, bias)
        y = torch.zeros(n, h, w, c, dtype=torch.float32, device=DEV)
        ops.conv2d(d, x, wp, bp, y)
        torch.cuda.synchronize()
        return y

    y1, y2, y12 = run(x1, 1), run(x2, 1), run((x1.float() + x2.float()).half(), 1)
    lin = (y12 - (y1 + y2 - bias)).abs().max().item()
    assert lin <= 2e-5 * y12.abs().max().item(), lin
    for mode in (4, 3):                 # CTA pairs with two clusters per SM pair; 256-row tiles
        assert torch.equal(run(x1, mode), y1), mode


def test_every_tuned_layer_variant_computes_what_the_default_kernel_computes(lib):
    """data/tuned_layers.json picks a kernel variant (N-tile width, CTA pair, patch kernel) per convolution shape of the BASELINE
    configurations at their FULL sizes.  Every entry: the variant's output equals the library's own choice on the same seeded
    operands up to the accumulation order (2e-3 of the output range; a wrong tiling would be off by O(1))."""
    import json
    from vehicle_counting_b200 import ops, _lib as L
    layers = json.load(open(os.path.join(ROOT, "vehicle_counting_b200", "data", "tuned_layers.json")))["layers"]
    checked = 0
    for key, v in sorted(layers.items()):
        k, s, cin, cout, res, hw, lg = key.split(",")
        k, s, cin, cout, res, n = int(k), int(s), int(cin), int(cout), int(res), 2 ** int(lg)
        h, w = (int(t) for t in hw.split("x"))
        f32 = cout == 255                                            # the Detect heads write fp32 logits
        cp = (cout + 7) // 8 * 8
        if n * h * w * max(cin, cp) * (4 if f32 else 2) > 3 << 30:
            continue
        g = torch.Generator().manual_seed(checked)
        x = torch.randn(n, h, w, cin, generator=g).half().to(DEV)
        wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
        b = (torch.randn(cout, generator=g) * 0.3).to(DEV)
        p = k // 2
        outs = []
        for bn, pair in ((0, 0), (v[0], v[1])):
            d = ops.make_conv_desc(n, h, w, cin, cout, k, s, p, cout_pitch=cp, act=L.ACT_NONE if f32 else L.ACT_SILU,
                                   res_mode=L.RES_AFTER_ACT if res else L.RES_NONE, res_pitch=cp if res else 0,
                                   out_dtype=L.F32 if f32 else L.F16, block_n=bn, cta_pair=pair)
            ho, wo = ops.conv_out_hw(d)
            wp, bp = ops.pack_conv_weights(d, wt, b)
            y = torch.zeros(n, ho, wo, cp, dtype=torch.float32 if f32 else torch.float16, device=DEV)
            r = (torch.randn(n, ho, wo, cp, generator=torch.Generator().manual_seed(7)).half().to(DEV)) if res else None
            ops.conv2d(d, x, wp, bp, y, residual=r)
            outs.append(y[..., :cout].float())
            del y, wp, bp, r
        torch.cuda.synchronize()
        assert tuple(L.last_fault()) == (0, 0, 0, 0), key
        err = (outs[0] - outs[1]).abs().max().item()
        assert err <= 2e-3 * max(1.0, outs[0].abs().max().item()), (key, v, err)
        checked += 1
        del outs, x
    assert checked >= 150
