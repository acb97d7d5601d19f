"""Where the host-side time of the numpy-in / numpy-out surface goes (development aid): python tools/e2e_breakdown.py"""
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
frames = [rng.integers(0, 256, (S, S, 3), dtype=np.uint8) for _ in range(B)]
cfg = types.SimpleNamespace(model_name="yolov5m", min_iou=0.45, min_conf=0.25, max_det=300)
det = ImageDetect(types.SimpleNamespace(weight=None, mapping=None, mapping_dict=None), cfg)
ex = Extractor("synthetic", bn_mode="train")
wh = rng.uniform(32, 256, (B, 64, 2)); tl = rng.uniform(0, 1, (B, 64, 2)) * (S - wh)
boxes = [np.concatenate([tl[i], tl[i] + wh[i]], 1) for i in range(B)]
for _ in range(3):
    det.run({"imgs": frames}); ex.from_frames(frames, boxes)
torch.cuda.synchronize()


def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


net = det.model.model
eng = net._engine(B, S, S)
pinned = net._pinned[(B, S, S)]
print("ImageDetect.run            %.2f ms" % t(lambda: det.run({"imgs": frames})))
print("  detect_raw               %.2f ms" % t(lambda: net.detect_raw(frames)))
print("  copy_frames (pool)       %.2f ms" % t(lambda: hostcopy.copy_frames(pinned.numpy(), frames)))
print("  serial np.copyto         %.2f ms" % t(lambda: [np.copyto(pinned.numpy()[i], f) for i, f in enumerate(frames)]))
print("  upload_frames (chunked)  %.2f ms" % t(lambda: hostcopy.upload_frames(pinned, eng.frames, frames, eng.plan.stream)))
print("  H2D only                 %.2f ms" % t(lambda: eng.upload(pinned)))
print("  forward (graph)          %.2f ms" % t(lambda: eng.forward()))
print("  download                 %.2f ms" % t(lambda: eng.download()))
print("Extractor.from_frames      %.2f ms" % t(lambda: ex.from_frames(frames, boxes)))
re = ex.engine
dev = re.stage_frame_list(frames)
rois = np.zeros((B * 64, 5), np.int32); rois[:, 0] = np.repeat(np.arange(B), 64); rois[:, 3:] = 100
print("  stage_frame_list         %.2f ms" % t(lambda: re.stage_frame_list(frames)))
print("  run (graph)              %.2f ms" % t(lambda: re.run(dev, rois, seg_sizes=[64] * B)))
print("  download(4096)           %.2f ms" % t(lambda: re.download(4096)))
print("cpu count", os.cpu_count(), "torch threads", torch.get_num_threads())
