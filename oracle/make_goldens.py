"""ORACLE tooling: generate the committed golden vectors under tests/golden/ and the git-ignored
weight repack under oracle/_ref/.  Run in the build container (needs /root/reference):

    python -m oracle.make_goldens

Outputs
  tests/golden/reid_golden.npz     seeded BGR crops + embeddings computed by the REFERENCE's own
                                   Extractor/Net with the shipped ckpt.t7: as shipped (train-mode BN),
                                   and with .eval(); plus the reference preprocessing of the crops
  tests/golden/deepsort_golden.npz a short synthetic sequence through the REFERENCE DeepSort.update
                                   (integer output rows per frame) with the features it was fed
  tests/golden/yolo_golden.npz     oracle (restatement) outputs on a seeded 2-frame 64x96 clip for
                                   yolov5n: raw heads + post-NMS rows (guards the oracle against drift;
                                   upstream itself is not importable -> "parity unpinned")
  tests/golden/videotracker_golden.npz  rows of the REFERENCE's modules/track.py VideoTracker.run over a 10-frame sequence
  tests/golden/pipeline_golden.{npz,csv}  the CSV written by the REFERENCE's unmodified CountingPipeline.run on a synthetic FFV1 clip
                                   (detector = CPU oracle plugged in for networks.get_model) + the clip's base frame
  oracle/_ref/reid_ckpt.npz        fp32 repack of ckpt.t7's `net_dict` (derived artefact, git-ignored,
                                   travels to the GPU box with the working tree)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim, reid, yolov5  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_OUT = os.path.join(ROOT, "oracle", "_ref")


def reid_crops(seed=7):
    rng = np.random.default_rng(seed)
    sizes = [(80, 40), (33, 120), (50, 50), (200, 90), (17, 23), (64, 64), (5, 9), (130, 131)]
    crops = []
    for (h, w) in sizes:
        # smooth-ish structured content rather than pure noise: low-res noise upsampled + noise
        base = rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3)).astype(np.float32)
        up = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]
        im = np.clip(up + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)
        crops.append(im)
    return crops


def make_reid():
    ex = ref_shim.reference_extractor()
    crops = reid_crops()
    pre = ex._preprocess(crops).numpy()
    assert ex.net.training, "reference Extractor is expected to stay in train mode"
    f_train = ex(crops)                       # as shipped: batch statistics of THIS call
    f_train_first3 = ex(crops[:3])            # same crops, different call composition
    ex2 = ref_shim.reference_extractor()      # fresh copy: the train-mode calls above moved running stats
    ex2.net.eval()
    f_eval = ex2(crops)
    flat = np.concatenate([c.reshape(-1) for c in crops])
    shapes = np.array([c.shape[:2] for c in crops], np.int32)
    np.savez_compressed(os.path.join(GOLD, "reid_golden.npz"), crops_flat=flat, crop_shapes=shapes,
                        preprocessed=pre.astype(np.float32), feat_train=f_train, feat_train_first3=f_train_first3,
                        feat_eval=f_eval)
    sd = reid.load_state_dict(ref_shim.REID_CKPT)
    os.makedirs(REF_OUT, exist_ok=True)
    np.savez(os.path.join(REF_OUT, "reid_ckpt.npz"), **{k: v.numpy() for k, v in sd.items()})
    print("reid golden:", f_train.shape, f_eval.shape, "ckpt tensors:", len(sd))


def make_deepsort():
    """3 boxes drifting 3 px/frame over a textured frame, n_init=3 (cam_04 settings)."""
    rng = np.random.default_rng(3)
    H, W = 240, 320
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ds = ref_shim.reference_deepsort(max_dist=0.2, min_confidence=0.3, nms_max_overlap=0.5, max_iou_distance=0.7,
                                     max_age=70, n_init=3, nn_budget=100)
    feats_log, rows, boxes_log = [], [], []
    boxes0 = np.array([[20, 30, 80, 120], [150, 40, 230, 160], [100, 150, 160, 230]], np.float64)
    conf = np.array([0.9, 0.8, 0.7])
    T = 8
    for t in range(T):
        b = boxes0 + 3.0 * t
        crops = reid.get_crops(b, frame)
        feats_log.append(ds.extractor(crops))
        out = ds.update(b.copy(), conf.copy(), frame)
        out = np.asarray(out, dtype=np.int64).reshape(-1, 7) if len(out) else np.zeros((0, 7), np.int64)
        rows.append(out)
        boxes_log.append(b)
    np.savez_compressed(os.path.join(GOLD, "deepsort_golden.npz"), frame=frame, boxes=np.stack(boxes_log), conf=conf,
                        feats=np.stack(feats_log), row_counts=np.array([r.shape[0] for r in rows]),
                        rows=np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.int64))
    print("deepsort golden rows per frame:", [r.shape[0] for r in rows])


def make_yolo():
    torch.manual_seed(0)
    m = yolov5.build("yolov5n", seed=0, obj_bias=-1.0)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (64, 96, 3), dtype=np.uint8) for _ in range(2)]
    dets, pred, raw = yolov5.autoshape_forward(m, imgs, size=96, return_raw=True)
    np.savez_compressed(os.path.join(GOLD, "yolo_golden.npz"), imgs=np.stack(imgs),
                        raw0=raw[0].numpy(), raw1=raw[1].numpy(), raw2=raw[2].numpy(),
                        det_counts=np.array([d.shape[0] for d in dets]),
                        dets=torch.cat(dets, 0).numpy() if dets else np.zeros((0, 6), np.float32))
    print("yolo golden dets:", [d.shape[0] for d in dets])


def make_videotracker():
    """The REFERENCE's own VideoTracker.run (modules/track.py:8-70: one DeepSort per class, per-class fan-out, Extractor in
    train mode as shipped) on 10 frames: a textured 240x320 frame that scrolls 3 px per step with nine boxes of three classes
    riding on it.  Rows: x1, y1, x2, y2, track_id, label."""
    ref_shim.install()
    from modules.track import VideoTracker  # type: ignore
    rng = np.random.default_rng(21)
    H, W, T = 240, 320, 10
    base = rng.integers(0, 256, (H // 8, W // 8, 3)).astype(np.float32)
    frame = np.clip(np.kron(base, np.ones((8, 8, 1), np.float32)) + rng.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)
    cfg = {"MAX_DIST": 0.2, "MIN_CONFIDENCE": 0.25, "NMS_MAX_OVERLAP": 0.5, "MAX_IOU_DISTANCE": 0.6, "MAX_AGE": 30, "N_INIT": 3,
           "NN_BUDGET": 60}                                   # configs/cam_configs.yaml: cam_04
    nc = 3
    vt = VideoTracker(nc, {"tracking_config": cfg}, {"num_frames": T}, ref_shim.REID_CKPT)
    tl0 = np.array([[20, 30], [100, 20], [180, 40], [30, 130], [120, 120], [200, 140], [60, 70], [150, 80], [230, 30]], np.float64)
    size = np.array([[50, 70], [40, 60], [60, 50], [70, 80], [45, 45], [55, 75], [35, 50], [65, 40], [50, 60]], np.float64)
    labels = np.array([0, 0, 0, 0, 2, 2, 2, 1, 0])
    scores = np.linspace(0.9, 0.4, 9)
    boxes_log, rows, shifts = [], [], []
    for t in range(T):
        shift = 3 * t
        fr = np.roll(frame, shift, axis=1)
        boxes = np.concatenate([tl0 + np.array([shift, 0.5 * t]), size], 1)     # xywh top-left, as ImageDetect returns them
        out = vt.run(fr, boxes.copy(), labels.copy(), scores.copy())
        r = np.concatenate([np.asarray(out["boxes"], np.int64).reshape(-1, 4), np.asarray(out["tracks"], np.int64).reshape(-1, 1),
                            np.asarray(out["labels"], np.int64).reshape(-1, 1)], 1)
        rows.append(r); boxes_log.append(boxes); shifts.append(shift)
    np.savez_compressed(os.path.join(GOLD, "videotracker_golden.npz"), frame=frame, shift=np.array(shifts), boxes=np.stack(boxes_log),
                        labels=labels, scores=scores, num_classes=np.array(nc), row_counts=np.array([r.shape[0] for r in rows]),
                        rows=np.concatenate(rows, 0), **{"cfg_" + k: np.array(v) for k, v in cfg.items()})
    print("videotracker golden rows per frame:", [r.shape[0] for r in rows])


PIPE_CFG = dict(T=8, H=320, W=320, step=4, model="yolov5n", obj_bias=-6.0,
                tracking={"MAX_DIST": 0.2, "MIN_CONFIDENCE": 0.25, "NMS_MAX_OVERLAP": 0.5, "MAX_IOU_DISTANCE": 0.6, "MAX_AGE": 30, "N_INIT": 3,
                          "NN_BUDGET": 60})


def drop_small_boxes(dets, min_side: float = 4.0):
    """A seeded random network also emits sub-pixel boxes; their crops are empty and the reference's Extractor dies inside
    cv2.resize (it has no guard).  Both sides of the pipeline golden drop boxes with a side under `min_side` pixels."""
    out = []
    for d in dets:
        if d["bboxes"].size:
            k = (d["bboxes"][:, 2] >= min_side) & (d["bboxes"][:, 3] >= min_side)
            d = {"bboxes": d["bboxes"][k], "classes": d["classes"][k], "scores": d["scores"][k]}
            if not k.any():
                d = {"bboxes": np.array(()), "classes": np.array(()), "scores": np.array(())}
        out.append(d)
    return out


def pipeline_clip_frames(base: np.ndarray, T: int, step: int):
    """the synthetic clip: the textured base frame scrolling `step` pixels per frame (BGR, as cv2 stores them in the file)"""
    return [np.roll(base, step * t, axis=1) for t in range(T)]


def write_pipeline_inputs(dirname: str, base: np.ndarray, T: int, step: int):
    """cam_04.avi (FourCC FFV1: lossless through cv2) + cam_04.json (a zone covering the frame, two directions)"""
    import json
    import cv2
    os.makedirs(dirname, exist_ok=True)
    h, w = base.shape[:2]
    path = os.path.join(dirname, "cam_04.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 10, (w, h))
    assert vw.isOpened()
    for f in pipeline_clip_frames(base, T, step):
        vw.write(f)
    vw.release()
    zone = {"shapes": [{"label": "zone", "points": [[2.5, 2.5], [w - 2.5, 2.5], [w - 2.5, h - 2.5], [2.5, h - 2.5]]},
                       {"label": "direction01", "points": [[10.0, h / 2], [w - 10.0, h / 2]]},
                       {"label": "direction02", "points": [[w - 10.0, h / 2], [10.0, h / 2]]}]}
    with open(os.path.join(dirname, "cam_04.json"), "w") as fh:
        json.dump(zone, fh)
    return path


def make_pipeline():
    """J1: the UNMODIFIED reference driver modules/__init__.py:CountingPipeline.run -- its own VideoLoader (cv2 on a
    FFV1 cam_04.avi), ImageDetect, Detector, VideoTracker (80 DeepSort instances, shipped ckpt.t7, BatchNorm as shipped),
    VideoCounting and CSV writer -- with ONE substitution: `networks.get_model` returns the CPU oracle of the detector network
    (upstream YOLOv5 cannot be fetched here), seeded weights.  The CSV it writes is the golden the GPU mirror is compared with."""
    import tempfile
    import types
    import torch.nn as nn
    ref_shim.install()
    c = PIPE_CFG
    rng = np.random.default_rng(33)
    base_lo = rng.integers(0, 256, (c["H"] // 16, c["W"] // 16, 3)).astype(np.float32)
    base = np.clip(np.kron(base_lo, np.ones((16, 16, 1), np.float32)) + rng.normal(0, 8, (c["H"], c["W"], 3)), 0, 255).astype(np.uint8)
    tmp = tempfile.mkdtemp(prefix="vcb_pipeline_")
    clip = write_pipeline_inputs(tmp, base, c["T"], c["step"])
    model = yolov5.build(c["model"], seed=0, obj_bias=c["obj_bias"])

    class OracleBackbone(nn.Module):             # the surface networks/yolo.py:YoloBackbone offers the stages
        def __init__(self):
            super().__init__()
            self.net = model
            self.class_names = [f"class{i}" for i in range(80)]

        def detect(self, batch, device):
            return drop_small_boxes(yolov5.yolo_backbone_detect(self.net, batch, size=640, conf=0.25, iou=0.45, max_det=300))

    import networks  # type: ignore  (the reference's package)
    import modules.detect as ref_detect  # type: ignore
    ref_detect.get_model = lambda args, config: OracleBackbone()
    from modules import CountingPipeline  # type: ignore
    args = types.SimpleNamespace(weight="seeded", input_path=clip, output_path=os.path.join(tmp, "out"), mapping=None, mapping_dict=None)
    config = types.SimpleNamespace(model_name=c["model"], min_iou=0.45, min_conf=0.25, max_det=300, image_size=[640, 640], keep_ratio=True)
    cam_config = types.SimpleNamespace(zone_path=tmp, checkpoint=ref_shim.REID_CKPT, cam={"cam_04": {"tracking_config": c["tracking"]}})
    pipe = CountingPipeline(args, config, cam_config)
    try:
        pipe.run()
    except Exception as e:                       # the overlay video rendered AFTER the CSV needs drawing deps stubbed out here
        print("[make_pipeline] note: post-CSV overlay step failed:", type(e).__name__, str(e)[:120])
    csv_path = os.path.join(tmp, "out", "cam_04.csv")
    assert os.path.isfile(csv_path), "the reference driver did not reach the CSV writer"
    import pandas as pd
    df = pd.read_csv(csv_path)
    print("pipeline golden: rows", len(df), "tracks", df.track_id.nunique(), "labels", sorted(df.label.unique().tolist()))
    with open(csv_path) as fh:
        text = fh.read()
    with open(os.path.join(GOLD, "pipeline_golden.csv"), "w") as fh:
        fh.write(text)
    np.savez_compressed(os.path.join(GOLD, "pipeline_golden.npz"), base=base, T=np.array(c["T"]), step=np.array(c["step"]),
                        obj_bias=np.array(c["obj_bias"]), **{"cfg_" + k: np.array(v) for k, v in c["tracking"].items()})


def main():
    os.makedirs(GOLD, exist_ok=True)
    if not ref_shim.available():
        raise SystemExit("reference tree not available; goldens can only be regenerated in the build container")
    make_reid()
    make_deepsort()
    make_yolo()
    make_videotracker()
    make_pipeline()


if __name__ == "__main__":
    main()
