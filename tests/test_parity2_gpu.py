"""Round-2 parity cases the round-1 verdict asked for (all through the C ABI, checker = the CPU oracle / reference goldens):

* scale_coords + clip with gain != 1 and pad != 0 (1280x720 -> 384x640, device letterbox), and a mixed-size image list
  (AutoShape's shape1 = max over the batch, host letterbox) -- rows in ORIGINAL pixels against oracle.yolov5.autoshape_forward
* YOLOv5l heads at 384x640 and 736x1280 (BASELINE configs[4]) against the fp32 oracle on one frame
* the optional class filter applied BEFORE the max_det cut (upstream non_max_suppression order)
* Extractor.__call__ called directly (feature_extractor.py:42-47), train and eval BatchNorm; Extractor.from_frames against
  per-frame calls
* VideoTracker.run rows against a golden emitted by the REFERENCE's own VideoTracker (tests/golden/videotracker_golden.npz,
  oracle/make_goldens.py)
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CKPT_NPZ = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")


def _textured(rng, h, w, cell=16, noise=8.0):
    base = rng.integers(0, 256, ((h + cell - 1) // cell, (w + cell - 1) // cell, 3)).astype(np.float32)
    up = np.kron(base, np.ones((cell, cell, 1), np.float32))[:h, :w]
    return np.clip(up + rng.normal(0, noise, (h, w, 3)), 0, 255).astype(np.uint8)


def _iou_xywh(a, b):
    x1 = np.maximum(a[0], b[:, 0]); y1 = np.maximum(a[1], b[:, 1])
    x2 = np.minimum(a[0] + a[2], b[:, 0] + b[:, 2]); y2 = np.minimum(a[1] + a[3], b[:, 1] + b[:, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    return inter / (a[2] * a[3] + b[:, 2] * b[:, 3] - inter)


def _rows_agree(got, want, conf=0.25, min_rate=0.9):
    """every oracle row away from the confidence threshold has a same-class twin with IoU > 0.9 and a close score"""
    n_firm = n_hit = 0
    worst_px = 0.0
    for g, w in zip(got, want):
        wb, wl, ws = w["bboxes"], w["classes"], w["scores"]
        for i in range(len(ws)):
            if abs(ws[i] - conf) < 5e-3:
                continue
            n_firm += 1
            if g["bboxes"].size == 0:
                continue
            same = np.nonzero(g["classes"] == wl[i])[0]
            if len(same) == 0:
                continue
            iou = _iou_xywh(wb[i], g["bboxes"][same])
            j = int(np.argmax(iou))
            if iou[j] > 0.9 and abs(g["scores"][same[j]] - ws[i]) < 6e-3:
                n_hit += 1
                worst_px = max(worst_px, float(np.abs(g["bboxes"][same[j]] - wb[i]).max()))
    assert n_firm > 0, "the seeded network must detect something"
    assert n_hit >= min_rate * n_firm, (n_hit, n_firm)
    return n_firm, n_hit, worst_px


def _backbone(model, **kw):
    from vehicle_counting_b200.networks.yolo import YoloBackbone
    return YoloBackbone(None, 0.45, 0.25, 300, state_dict=model.state_dict(), **kw)


def _check_scale_coords_exact(net, imgs, size):
    """The engine's rows must be oracle.scale_coords (upstream scale_coords + clip_coords: (x - pad) / gain, clamp to the ORIGINAL
    frame) of the rows the oracle's greedy NMS keeps from the engine's own candidates -- identical inputs, so this isolates the
    rescaling arithmetic with its real gain / pad values.  Returns the number of rows checked."""
    from oracle import yolov5 as Y
    eng = next(iter(net._engines.values()))
    det = eng.det.cpu().numpy(); cnt = eng.det_count.cpu().numpy()
    n_cand = eng.cand_count.cpu().numpy()
    ci = eng.cand_index.cpu().numpy(); cs = eng.cand_score.cpu().numpy(); cb = eng.cand_box.cpu().numpy(); cc = eng.cand_cls.cpu().numpy()
    rows = 0
    for b, im in enumerate(imgs):
        k = n_cand[b]
        order = np.lexsort((ci[b, :k], -cs[b, :k]))
        boxes = cb[b, :k][order] + (cc[b, :k][order].astype(np.float32) * np.float32(eng.max_wh))[:, None]
        keep = Y.greedy_nms(boxes, np.arange(len(order), 0, -1, dtype=np.float32), eng.iou)[:eng.max_det]
        sel = order[keep]
        assert cnt[b] == len(sel), (b, cnt[b], len(sel))
        want = Y.scale_coords((eng.h, eng.w), torch.from_numpy(cb[b, sel].copy()), im.shape[:2]).numpy()
        np.testing.assert_allclose(det[b, :cnt[b], :4], want, rtol=0, atol=2e-3)        # fp32 (x - pad) / gain on both sides
        np.testing.assert_array_equal(det[b, :cnt[b], 4], cs[b, sel])
        h0, w0 = im.shape[:2]
        d = det[b, :cnt[b]]
        assert (d[:, [0, 2]] >= 0).all() and (d[:, [0, 2]] <= w0).all() and (d[:, [1, 3]] >= 0).all() and (d[:, [1, 3]] <= h0).all()
        rows += int(cnt[b])
    return rows


# End-to-end row agreement with the oracle on these textured frames is a RATE: the seeded random network fires hundreds of heavily
# overlapping boxes, so one IoU that lands on the other side of 0.45 re-routes a whole NMS cascade.  Yardstick measured on the CPU
# (no CUDA involved): the oracle's own fp16-storage twin agrees with the fp32 oracle on 80-87 % of the rows by the same criterion.
E2E_MIN_RATE = 0.7


def test_scale_coords_1280x720_device_letterbox_vs_oracle(lib):
    """gain = 0.5, pad = (0, 12): boxes come back in 1280x720 pixels, clipped to the frame (upstream scale_coords / clip_coords)."""
    from oracle import yolov5 as Y
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    rng = np.random.default_rng(11)
    imgs = [_textured(rng, 720, 1280) for _ in range(2)]
    model = Y.build("yolov5n", seed=0, obj_bias=-5.0)
    net = _backbone(model)
    got = net.detect({"imgs": imgs})
    assert (384, 640) in [(k[1], k[2]) for k in net._engines], "1280x720 must run at the reference's 384x640 inference shape"
    assert _check_scale_coords_exact(net, imgs, 640) > 50
    want = Y.yolo_backbone_detect(Y.fp16_storage_twin(model), {"imgs": imgs}, size=640)
    _rows_agree(got, want, min_rate=E2E_MIN_RATE)


def test_mixed_size_image_list_vs_oracle(lib):
    """AutoShape: shape1 = max over the batch of size * (h, w) / max(h, w), rounded up to the stride; every image gets its own
    gain / pad in scale_coords (networks/yolo.py:68-99 passes the list straight through)."""
    from oracle import yolov5 as Y
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    rng = np.random.default_rng(12)
    imgs = [_textured(rng, 480, 640), _textured(rng, 360, 640), _textured(rng, 640, 400), _textured(rng, 200, 300)]
    model = Y.build("yolov5n", seed=0, obj_bias=-5.0)
    net = _backbone(model, size=320)
    got = net.detect({"imgs": imgs})
    assert _check_scale_coords_exact(net, imgs, 320) > 50
    want = Y.yolo_backbone_detect(Y.fp16_storage_twin(model), {"imgs": imgs}, size=320)
    _rows_agree(got, want, min_rate=E2E_MIN_RATE)


@pytest.mark.parametrize("hw", [(384, 640), (736, 1280)])
def test_yolov5l_heads_vs_oracle_one_frame(lib, hw):
    """BASELINE configs[4]: YOLOv5l at the reference's inference shape for 1280x720 frames (size=640 -> 384x640) and at size=1280
    (736x1280): head tensors of one letterboxed frame against the fp32 oracle and its fp16-storage twin."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import YoloEngine
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    model = Y.build("yolov5l", seed=0, obj_bias=-3.0)
    frame = _textured(np.random.default_rng(13), 720, 1280)
    size = max(hw)
    _, _, raw_ref = Y.autoshape_forward(model, [frame], size=size, return_raw=True)
    _, _, raw_twin = Y.autoshape_forward(Y.fp16_storage_twin(model), [frame], size=size, return_raw=True)
    lb = Y.letterbox(frame, hw)                            # what AutoShape feeds the network (cv2 resize + pad 114)
    assert lb.shape[:2] == hw
    eng = YoloEngine(model.state_dict(), 1, hw[0], hw[1], model_name="yolov5l")
    eng.upload(torch.from_numpy(lb[None]).pin_memory()); eng.forward()
    eng.download()
    for li in range(3):
        got = eng.logits[li].float().cpu()[..., :3 * eng.no].permute(0, 3, 1, 2)
        rel32 = ((got - raw_ref[li]).norm() / raw_ref[li].norm()).item()
        rel16 = ((got - raw_twin[li]).norm() / raw_twin[li].norm()).item()
        assert rel32 < 7e-3 and rel16 < 3e-3, (hw, li, rel32, rel16)


def test_class_filter_runs_before_the_max_det_cut(lib):
    """upstream non_max_suppression filters `classes` before max_nms / NMS / max_det: with max_det small, the rows must be the top
    rows OF THE KEPT CLASSES, not the kept-class subset of the overall top rows."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import YoloEngine
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    model = Y.build("yolov5n", seed=0, obj_bias=-1.0)
    rng = np.random.default_rng(14)
    imgs = [_textured(rng, 160, 160) for _ in range(2)]
    twin = Y.fp16_storage_twin(model)
    all_rows = Y.autoshape_forward(twin, imgs, size=160, max_det=300)
    present = sorted({int(c) for d in all_rows for c in d[:, 5].tolist()})
    assert len(present) >= 4
    keep = present[1::2]                                  # every other class that actually fires
    want = Y.autoshape_forward(twin, imgs, size=160, classes=keep, max_det=5)
    eng = YoloEngine(model.state_dict(), 2, 160, 160, model_name="yolov5n", max_det=5, classes=keep)
    eng.upload(torch.from_numpy(np.stack(imgs)).pin_memory()); eng.forward()
    det, cnt = eng.download()
    for b in range(2):
        assert set(det[b, :cnt[b], 5].astype(int).tolist()) <= set(keep)
        assert cnt[b] == want[b].shape[0], (cnt[b], want[b].shape[0])
        if cnt[b]:
            np.testing.assert_allclose(det[b, :cnt[b], 4], want[b][:, 4].numpy(), atol=6e-3)
            gaps = np.abs(np.diff(want[b][:, 4].numpy()))
            if gaps.size == 0 or gaps.min() > 1.2e-2:          # no near-tie in the oracle's ranking: same classes in the same order
                assert det[b, :cnt[b], 5].astype(int).tolist() == want[b][:, 5].int().tolist()


def _reid_path_and_sd():
    from oracle import reid as R
    if os.path.isfile(CKPT_NPZ):
        return CKPT_NPZ, R.load_state_dict(CKPT_NPZ)
    return "synthetic", None


@pytest.mark.parametrize("bn_mode", ["train", "eval"])
def test_extractor_call_direct(lib, bn_mode):
    """Extractor(model_path, use_cuda)(im_crops) -> float32 [n, 512] (feature_extractor.py:42-47), crops of different sizes."""
    from oracle import reid as R
    from vehicle_counting_b200.networks.deepsort.deep_sort import Extractor
    from vehicle_counting_b200.weights import synth_reid_state_dict
    path, sd = _reid_path_and_sd()
    if sd is None:
        sd = synth_reid_state_dict(0)
    rng = np.random.default_rng(15)
    crops = [_textured(rng, h, w, cell=8) for h, w in [(80, 40), (33, 120), (50, 50), (200, 90), (17, 23), (64, 64), (130, 131)]]
    ex = Extractor(path, use_cuda=True, bn_mode=bn_mode)
    for rep in range(3):                                   # repeated calls with different compositions reuse the same plans
        sub = crops[:len(crops) - rep]
        got = ex(sub)
        ref = R.extract(sd, sub, bn_mode)
        assert got.dtype == np.float32 and got.shape == (len(sub), 512)
        assert np.abs(got - ref).max() < (4e-3 if bn_mode == "train" else 1.5e-3), (bn_mode, rep, np.abs(got - ref).max())
    n_plans = len(ex.engine._plans)
    for _ in range(4):
        ex(crops)
    assert len(ex.engine._plans) == n_plans                # no plan / graph is built per call (unbounded growth)
    assert ex([]).shape == (0, 512)


@pytest.mark.parametrize("bn_mode", ["train", "eval"])
def test_from_frames_equals_per_frame_calls(lib, bn_mode):
    """the batched multi-frame feature call = one reference call per frame (one BatchNorm segment per frame)"""
    from oracle import reid as R
    from vehicle_counting_b200.networks.deepsort.deep_sort import Extractor
    from vehicle_counting_b200.weights import synth_reid_state_dict
    path, sd = _reid_path_and_sd()
    if sd is None:
        sd = synth_reid_state_dict(0)
    rng = np.random.default_rng(16)
    frames = [_textured(rng, 240, 320, cell=8) for _ in range(3)]
    boxes = []
    for k in (5, 1, 7):
        wh = rng.uniform(20, 120, (k, 2)); tl = rng.uniform(0, 1, (k, 2)) * (np.array([320, 240]) - wh)
        boxes.append(np.concatenate([tl, tl + wh], 1))
    ex = Extractor(path, use_cuda=True, bn_mode=bn_mode)
    got = ex.from_frames(frames, boxes)
    for f, b, g in zip(frames, boxes, got):
        ref = R.extract(sd, R.get_crops(b, f), bn_mode)
        tol = 1.5e-3 if bn_mode == "eval" else (4e-3 if len(b) > 1 else 2e-2)     # 1-crop segment: 16 values per channel at the last stage
        assert g.shape == ref.shape and np.abs(g - ref).max() < tol, (bn_mode, len(b), np.abs(g - ref).max())


def test_video_tracker_rows_match_the_reference_golden(lib):
    """rows emitted by the REFERENCE's own VideoTracker.run (modules/track.py:30-70, CPU, shipped ckpt.t7, BatchNorm as shipped)
    on a drifting-box sequence; the GPU mirror must give the same track ids, labels and integer boxes."""
    path = os.path.join(GOLD, "videotracker_golden.npz")
    if not os.path.isfile(path) or not os.path.isfile(CKPT_NPZ):
        pytest.skip("golden or shipped ReID weights not present")
    from vehicle_counting_b200.modules import VideoTracker
    z = np.load(path)
    cam = {"tracking_config": {k: z["cfg_" + k].item() for k in ("MAX_DIST", "MIN_CONFIDENCE", "NMS_MAX_OVERLAP", "MAX_IOU_DISTANCE",
                                                                  "MAX_AGE", "N_INIT", "NN_BUDGET")}}
    nc = int(z["num_classes"])
    vt = VideoTracker(nc, cam, {"num_frames": int(z["boxes"].shape[0])}, CKPT_NPZ, bn_mode="train")
    off = 0
    total = 0
    for t in range(z["boxes"].shape[0]):
        frame = np.roll(z["frame"], int(z["shift"][t]), axis=1)
        out = vt.run(frame, z["boxes"][t].copy(), z["labels"].copy(), z["scores"].copy())
        k = int(z["row_counts"][t])
        want = z["rows"][off:off + k]                       # columns: x1, y1, x2, y2, track_id, label
        off += k
        assert [int(v) for v in out["tracks"]] == want[:, 4].tolist(), t
        assert [int(v) for v in out["labels"]] == want[:, 5].tolist(), t
        np.testing.assert_array_equal(np.asarray(out["boxes"]).reshape(-1, 4), want[:, :4])
        total += k
    assert total > 0


def test_v5_checkpoint_focus_spp_heads_vs_oracle(lib):
    """north_star names the <= v5.0 building blocks (Focus stem, SPP): a v5.0-layout state_dict (model.0.conv.conv.*, SPP at
    model.8, nine C3 repeats at P3) runs through the same kernels -- Focus == the space-to-depth stem with a channel permutation,
    SPP(5, 9, 13) == the SPPF cascade -- and matches the oracle's v5.0 graph."""
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import YoloEngine
    torch.set_num_threads(max(os.cpu_count() or 1, 1))
    model = Y.build("yolov5s", seed=0, obj_bias=-2.0, version="v5")
    rng = np.random.default_rng(18)
    imgs = [rng.integers(0, 256, (160, 224, 3), dtype=np.uint8) for _ in range(2)]
    _, _, raw_ref = Y.autoshape_forward(model, imgs, size=224, return_raw=True)
    dets_twin, _, raw_twin = Y.autoshape_forward(Y.fp16_storage_twin(model), imgs, size=224, return_raw=True)
    eng = YoloEngine(model.state_dict(), 2, 160, 224)
    assert eng.version == "v5" and eng.name == "yolov5s"
    eng.upload(torch.from_numpy(np.stack(imgs)).pin_memory()); eng.forward()
    det, cnt = eng.download()
    for li in range(3):
        got = eng.logits[li].float().cpu()[..., :3 * eng.no].permute(0, 3, 1, 2)
        rel32 = ((got - raw_ref[li]).norm() / raw_ref[li].norm()).item()
        rel16 = ((got - raw_twin[li]).norm() / raw_twin[li].norm()).item()
        assert rel32 < 7e-3 and rel16 < 3e-3, (li, rel32, rel16)
    n_ref = sum(d.shape[0] for d in dets_twin)
    assert abs(int(cnt.sum()) - n_ref) <= max(3, n_ref // 10), (cnt.tolist(), n_ref)
