"""Tracked-pipeline timing through the reference-shaped call surface (development aid; run under gpurun):
ImageDetect-like detection on 1280x720 frames (YoloBackbone.detect, device letterbox to 384x640) in batches, then per frame
VideoTracker.run (ONE ReID pass per frame + the host association step) on 64 synthetic moving boxes of 3 classes.
Synthetic detector weights produce noise detections, so the tracker is fed the synthetic boxes (as bench.py feeds synthetic
ROIs); the detector still runs on every frame.  Prints one JSON line with frames/s and the split.

    python tools/pipeline_demo.py --model yolov5l --frames 96 --batch 32 --bn eval
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolov5l")
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tracks", type=int, default=64)
    ap.add_argument("--bn", default="eval", choices=["eval", "train"])
    ap.add_argument("--size", type=int, default=640, help="AutoShape size: 640 -> 384x640 (the reference), 1280 -> 736x1280")
    a = ap.parse_args()
    from vehicle_counting_b200.modules import VideoTracker
    from vehicle_counting_b200.networks.yolo import YoloBackbone
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    rng = np.random.default_rng(0)
    H, W, n = 720, 1280, a.tracks
    net = YoloBackbone(None, 0.45, 0.25, 300, state_dict=synth_yolov5_state_dict(a.model, seed=0, obj_bias=-3.0), size=a.size)
    cam = {"tracking_config": {"MAX_DIST": 0.3, "MIN_CONFIDENCE": 0.3, "NMS_MAX_OVERLAP": 0.5, "MAX_IOU_DISTANCE": 0.7, "MAX_AGE": 30,
                               "N_INIT": 3, "NN_BUDGET": 50}}
    vt = VideoTracker(3, cam, {"num_frames": a.frames}, "synthetic", bn_mode=a.bn)
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(a.batch)]
    pos = rng.uniform([0, 0], [W - 260, H - 260], (n, 2)); vel = rng.uniform(-3, 3, (n, 2)); size = rng.uniform(32, 250, (n, 2))
    labels = rng.integers(0, 3, n)
    scores = rng.uniform(0.4, 0.95, n)

    def run(nframes, timed):
        t_det = t_trk = 0.0
        rows = 0
        for f0 in range(0, nframes, a.batch):
            t0 = time.perf_counter()
            net.detect({"imgs": frames})                       # one batch of frames through the detector adapter
            t1 = time.perf_counter()
            for k in range(a.batch):
                t = f0 + k
                p = np.clip(pos + vel * t, 0, [W - 260, H - 260])
                boxes = np.concatenate([p, size], 1)            # xywh top-left, as ImageDetect returns them
                out = vt.run(frames[k], boxes, labels, scores)
                rows += len(out["tracks"])
            t2 = time.perf_counter()
            t_det += t1 - t0; t_trk += t2 - t1
        return t_det, t_trk, rows

    run(a.batch, False)                                         # warm-up (plans, graphs)
    torch.cuda.synchronize()
    t_det, t_trk, rows = run(a.frames, True)
    torch.cuda.synchronize()
    tot = t_det + t_trk
    print(json.dumps({"pipeline": f"{a.model} 1280x720 -> {'384x640' if a.size == 640 else '736x1280'} detect (batch {a.batch}) + VideoTracker ({n} boxes, 3 classes, bn {a.bn})",
                      "frames": a.frames, "fps": a.frames / tot, "detect_ms_per_frame": 1e3 * t_det / a.frames,
                      "track_ms_per_frame": 1e3 * t_trk / a.frames, "rows": rows}))


if __name__ == "__main__":
    main()
