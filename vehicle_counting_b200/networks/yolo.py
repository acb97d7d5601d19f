"""`get_model` / `YoloBackbone` with the reference's signatures (/root/reference/networks/yolo.py:11-99),
backed by the CUDA engine instead of `torch.hub.load('ultralytics/yolov5', ...)`.

`YoloBackbone.detect(batch, device)` keeps the reference contract: `batch['imgs']` is a list of HWC uint8
RGB arrays (sizes may differ) and the result is one dict per image with float64 `bboxes`
(x_min, y_min, w, h in original-image pixels), integer `classes` and `scores`, rows sorted by score,
at most `max_det` rows, three empty arrays when nothing is found.  The host half of upstream's AutoShape
(shape bookkeeping, cv2 letterbox) is restated here; everything after the uint8 batch reaches the device
runs in libvcb200's kernels.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from ..engine import YoloEngine
from ..hostcopy import copy_frames, upload_frames
from ..weights import load_yolov5_checkpoint, synth_yolov5_state_dict

COCO_NAMES = ['person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light',
              'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant',
              'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard',
              'sports ball', 'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle',
              'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli',
              'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair', 'couch', 'potted plant', 'bed', 'dining table', 'toilet',
              'tv', 'laptop', 'mouse', 'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink',
              'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear', 'hair drier', 'toothbrush']


def get_model(args, config):
    """networks/yolo.py:11-34.  `args.weight` must name a v6.0 state_dict file; the reference's download
    branch (`args.weight is None`) needs a network and is replaced by an explicit error unless
    VCB_SYNTH_WEIGHTS=1 asks for the seeded synthetic stand-in (benchmarks, tests)."""
    filter_classes = None if not getattr(args, "mapping_dict", None) else args.mapping_dict.keys()
    weight = getattr(args, "weight", None)
    if weight is None:
        if os.environ.get("VCB_SYNTH_WEIGHTS", "0") != "1":
            raise FileNotFoundError(
                f"no --weight given and pretrained '{config.model_name}' cannot be downloaded here "
                "(set VCB_SYNTH_WEIGHTS=1 to run with seeded synthetic weights)")
        weight = f"synthetic:{config.model_name}"
    return YoloBackbone(weight=weight, min_iou=config.min_iou, min_conf=config.min_conf, max_det=config.max_det,
                        filter_classes=filter_classes)


class BaseBackbone(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, batch):
        pass

    def detect(self, batch):
        pass


def _make_divisible(x: float, d: int) -> int:
    return int(math.ceil(x / d) * d)


def _letterbox(im: np.ndarray, new_shape: Tuple[int, int]) -> np.ndarray:
    """upstream utils/augmentations.py letterbox(auto=False, scaleup=True), on the host like upstream."""
    import cv2
    h, w = im.shape[:2]
    r = min(new_shape[0] / h, new_shape[1] / w)
    new_unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    if (w, h) != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    if top or bottom or left or right:
        im = cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    return im


def cv2_linear_table(dst_n: int, src_n: int, vertical: bool) -> np.ndarray:
    """The per-column (per-row) table of cv2.resize(INTER_LINEAR) for 8-bit images [OpenCV imgproc/resize.cpp, 11-bit fixed point]:
    int32 [dst_n, 4] = source index 0, source index 1, weight 0, weight 1 (x 2048).  Restated operation by operation: the scale is
    1 / (dst / src) in double, the source coordinate is rounded to float32 before the floor, the weights are float32 products
    rounded half-to-even; columns clamp the coordinate AND zero the fraction at the borders, rows only clamp the two row indices
    (so a border row blends the same source row with both weights).  tests/test_geometry_cpu.py checks the table against
    cv2.resize on 50 size pairs; csrc/pointwise.cu letterbox_bilinear_kernel consumes it."""
    scale = 1.0 / (float(dst_n) / float(src_n))
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo, hi = s < 0, s >= src_n - 1
        f[lo] = 0; s[lo] = 0
        f[hi] = 0; s[hi] = src_n - 1
    one, k = np.float32(1.0), np.float32(2048.0)
    a0 = np.rint(((one - f) * k).astype(np.float32))
    a1 = np.rint((f * k).astype(np.float32))
    return np.stack([np.clip(s, 0, src_n - 1), np.clip(s + 1, 0, src_n - 1), a0, a1], 1).astype(np.int32)


class YoloBackbone(BaseBackbone):
    def __init__(self, weight, min_iou, min_conf, max_det, filter_classes=None, size: int = 640, device: Optional[str] = None,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, class_names: Optional[Sequence[str]] = None, **kwargs):
        super().__init__(**kwargs)
        if state_dict is None:
            if isinstance(weight, str) and weight.startswith("synthetic:"):
                state_dict = synth_yolov5_state_dict(weight.split(":", 1)[1], seed=0)
            else:
                state_dict, names = load_yolov5_checkpoint(weight)      # custom models carry their own class names
                if class_names is None:
                    class_names = names
        self._sd = state_dict
        self.size = size                       # AutoShape's `size=` (the reference always uses the default 640)
        self.conf, self.iou, self.max_det = min_conf, min_iou, max_det
        self.classes = list(filter_classes) if filter_classes is not None else None
        self.multi_label = False
        nc = state_dict["model.24.m.0.weight"].shape[0] // 3 - 5
        if class_names is None:
            class_names = COCO_NAMES if nc == 80 else [f"class{i}" for i in range(nc)]
        self.class_names = list(class_names)
        self._device = device
        self._engines: Dict[Tuple[int, int, int], YoloEngine] = {}
        self._pinned: Dict[tuple, torch.Tensor] = {}
        self._raw_dev: Dict[tuple, torch.Tensor] = {}
        self._lb_tables: Dict[tuple, Tuple[torch.Tensor, torch.Tensor, tuple]] = {}      # (h0, w0, h1, w1) -> device tables + geometry
        # a parameter so that `.parameters()` / `.to()` behave like the reference module (detect.py:27-28)
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)

    def _engine(self, b: int, h: int, w: int) -> YoloEngine:
        key = (b, h, w)
        if key not in self._engines:
            dev = self._device or (f"cuda:{torch.cuda.current_device()}")
            self._engines[key] = YoloEngine(self._sd, b, h, w, device=dev, conf=self.conf, iou=self.iou, max_det=self.max_det,
                                            classes=self.classes)
            self._pinned[key] = torch.empty(b, h, w, 3, dtype=torch.uint8).pin_memory()
        return self._engines[key]

    @torch.no_grad()
    def detect_raw(self, imgs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        """-> (det float32 [B, max_det, 6] xyxy/conf/cls in original pixels, counts int32 [B])."""
        shape0 = [im.shape[:2] for im in imgs]
        shape1 = [0.0, 0.0]
        for (h, w) in shape0:                                   # AutoShape: g = size / max(h, w)
            g = self.size / max(h, w)
            shape1 = [max(shape1[0], h * g), max(shape1[1], w * g)]
        h1, w1 = (_make_divisible(v, 32) for v in shape1)
        eng = self._engine(len(imgs), h1, w1)
        eng.set_scale(shape0)
        h0, w0 = shape0[0]
        r = min(h1 / h0, w1 / w0)
        if all(s == shape0[0] for s in shape0) and r == 0.5 and h0 % 2 == 0 and w0 % 2 == 0:
            # exact 2x reduction (e.g. 1280x720 -> 640x360 inside 384x640): raw frames go up, the letterbox runs on the device
            # (vcb_letterbox_half_u8: bit-identical to the cv2 path below for this ratio)
            key = ("raw", len(imgs), h0, w0)
            if key not in self._pinned:
                self._pinned[key] = torch.empty(len(imgs), h0, w0, 3, dtype=torch.uint8).pin_memory()
                self._raw_dev[key] = torch.empty(len(imgs), h0, w0, 3, dtype=torch.uint8, device=eng.frames.device)
            host, raw = self._pinned[key], self._raw_dev[key]
            dh, dw = (h1 - h0 // 2) / 2, (w1 - w0 // 2) / 2
            top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
            upload_frames(host, raw, imgs, eng.plan.stream)
            ops.letterbox_half(raw, len(imgs), h0, w0, eng.frames, h1, w1, top, left, 114, stream=eng.plan.stream)
        elif all(s == shape0[0] for s in shape0) and (h0, w0) != (h1, w1) and os.environ.get("VCB_DEVICE_LETTERBOX", "1") != "0":
            # same-size frames at any other ratio: raw frames go up, cv2's fixed-point bilinear + the 114 border run on the device
            # (vcb_letterbox_bilinear_u8: bit-identical to the host cv2 path below)
            key = ("raw", len(imgs), h0, w0)
            if key not in self._pinned:
                self._pinned[key] = torch.empty(len(imgs), h0, w0, 3, dtype=torch.uint8).pin_memory()
                self._raw_dev[key] = torch.empty(len(imgs), h0, w0, 3, dtype=torch.uint8, device=eng.frames.device)
            tkey = (h0, w0, h1, w1)
            if tkey not in self._lb_tables:
                nw, nh = int(round(w0 * r)), int(round(h0 * r))
                dh, dw = (h1 - nh) / 2, (w1 - nw) / 2
                geo = (int(round(dh - 0.1)), int(round(dw - 0.1)), nh, nw)
                xt = torch.from_numpy(cv2_linear_table(nw, w0, False)).to(eng.frames.device)
                yt = torch.from_numpy(cv2_linear_table(nh, h0, True)).to(eng.frames.device)
                self._lb_tables[tkey] = (xt, yt, geo)
            xt, yt, (top, left, nh, nw) = self._lb_tables[tkey]
            host, raw = self._pinned[key], self._raw_dev[key]
            upload_frames(host, raw, imgs, eng.plan.stream)
            ops.letterbox_bilinear(raw, len(imgs), h0, w0, eng.frames, h1, w1, top, left, nh, nw, xt, yt, 114, stream=eng.plan.stream)
        else:
            host = self._pinned[(len(imgs), h1, w1)]
            if all(im.shape[:2] == (h1, w1) for im in imgs):
                upload_frames(host, eng.frames, imgs, eng.plan.stream)       # gather of chunk i+1 overlaps the H2D of chunk i
            else:
                hv = host.numpy()
                for i, im in enumerate(imgs):
                    hv[i] = im if im.shape[:2] == (h1, w1) else _letterbox(im, (h1, w1))
                eng.upload(host)
        eng.forward()
        # the optional class filter (yolo.py:64) runs inside the decode kernel, before the max_nms / max_det cuts
        return eng.download()

    def detect(self, batch, device=None):
        """networks/yolo.py:68-99"""
        det, cnt = self.detect_raw(batch["imgs"])
        # the reference round-trips through DataFrame.to_json (10 decimal digits) before np.array; one vectorised pass per batch
        nmax = int(cnt.max()) if len(cnt) else 0
        d = np.round(det[:, :nmax].astype(np.float64), 10)
        xywh = np.concatenate([d[..., :2], d[..., 2:4] - d[..., :2]], -1)
        cls = det[:, :nmax, 5].astype(np.int64)
        out = []
        for b in range(len(batch["imgs"])):
            n = int(cnt[b])
            if n > 0:
                out.append({"bboxes": xywh[b, :n].copy(), "classes": cls[b, :n].copy(), "scores": d[b, :n, 4].copy()})
            else:
                out.append({"bboxes": np.array(()), "classes": np.array(()), "scores": np.array(())})
        return out
