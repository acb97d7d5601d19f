"""Finer host-side timing of the numpy surface (development aid): where ImageDetect.run / Extractor.from_frames spend their time."""
import os, sys, time, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VCB_SYNTH_WEIGHTS"] = "1"
os.environ["VCB_REID_CAPACITY"] = "4096"
from vehicle_counting_b200.modules import ImageDetect
from vehicle_counting_b200.networks.deepsort.deep_sort import Extractor
from vehicle_counting_b200 import hostcopy

B, S = 64, 640
rng = np.random.default_rng(0)
host_pool = [torch.from_numpy(rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)).pin_memory() for _ in range(2)]
frames_sets = [[hp.numpy()[k] for k in range(B)] for hp in host_pool]
frames = frames_sets[0]
cfg = types.SimpleNamespace(model_name="yolov5m", min_iou=0.45, min_conf=0.25, max_det=300)
det = ImageDetect(types.SimpleNamespace(weight=None, mapping=None, mapping_dict=None), cfg)
ex = Extractor("synthetic", bn_mode="train")
wh = rng.uniform(32, 256, (B, 64, 2)); tl = rng.uniform(0, 1, (B, 64, 2)) * (S - wh)
boxes = [np.concatenate([tl[i], tl[i] + wh[i]], 1) for i in range(B)]
for i in range(3):
    det.run({"imgs": frames_sets[i % 2]}); ex.from_frames(frames_sets[i % 2], boxes)
torch.cuda.synchronize()
net = det.model.model
eng = net._engine(B, S, S)
pinned = net._pinned[(B, S, S)]
re = ex.engine


def lap(label, fn, n=10, sync=True):
    ts = []
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(i)
        if sync:
            torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{label:48s} median {np.median(ts):7.2f} ms   min {min(ts):7.2f}")


lap("ImageDetect.run", lambda i: det.run({"imgs": frames_sets[i % 2]}))
lap("  detect_raw", lambda i: net.detect_raw(frames_sets[i % 2]))
lap("  detect (raw + conversion)", lambda i: net.detect({"imgs": frames_sets[i % 2]}))
lap("  set_scale", lambda i: eng.set_scale([(S, S)] * B))
lap("  copy_frames 64 (pool)", lambda i: hostcopy.copy_frames(pinned.numpy(), frames_sets[i % 2]), sync=False)
lap("  copy_frames 32 (pool)", lambda i: hostcopy.copy_frames(pinned.numpy()[:32], frames_sets[i % 2][:32]), sync=False)
lap("  upload_frames (sync at end)", lambda i: hostcopy.upload_frames(pinned, eng.frames, frames_sets[i % 2], eng.plan.stream))
lap("  H2D only", lambda i: eng.upload(pinned))
lap("  forward", lambda i: eng.forward())
lap("  download", lambda i: eng.download())
lap("Extractor.from_frames", lambda i: ex.from_frames(frames_sets[i % 2], boxes))
dev = re.stage_frame_list(frames)
rois = np.zeros((B * 64, 5), np.int32); rois[:, 0] = np.repeat(np.arange(B), 64); rois[:, 3:] = 100
lap("  stage_frame_list", lambda i: re.stage_frame_list(frames_sets[i % 2]))
lap("  run", lambda i: re.run(dev, rois, seg_sizes=[64] * B))
lap("  download(4096)", lambda i: re.download(4096))


def roi_math(i):
    b = np.concatenate([np.asarray(x, np.float64).reshape(-1, 4) for x in boxes], 0)
    bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    cx, cy = b[:, 0] + bw / 2, b[:, 1] + bh / 2
    x1 = np.maximum(np.trunc(cx - bw / 2), 0); x2 = np.minimum(np.trunc(cx + bw / 2), S - 1)
    y1 = np.maximum(np.trunc(cy - bh / 2), 0); y2 = np.minimum(np.trunc(cy + bh / 2), S - 1)
    return np.stack([np.repeat(np.arange(B), 64), x1, y1, x2, y2], 1).astype(np.int32)


lap("  roi rule (numpy)", roi_math, sync=False)
print("cpu count", os.cpu_count(), "torch threads", torch.get_num_threads())
