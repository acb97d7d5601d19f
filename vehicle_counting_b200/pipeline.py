"""Batch pipeline over host frames: upload -> detect (+NMS) -> ROI crops -> ReID embeddings -> download.

This is the host-side driver a caller uses when frames and boxes live in host memory (as in the reference's
`CountingPipeline.run` loop, /root/reference/modules/__init__.py:54-84): uploads of batch i+1 run on a copy stream
while batch i computes, results come back through pinned buffers.  All arithmetic is libvcb200 (engine.py plans).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .engine import ReidEngine, YoloEngine


class FramePipeline:
    """Double-buffered host->device->host pipeline around a YoloEngine and a ReidEngine.

    submit(frames_pinned, rois) enqueues one batch and returns immediately; collect() returns the results of the
    oldest batch in flight: (det [B, max_det, 6] float32, det_count [B] int32, features [n_rois, 512] float32).
    ROIs are host int32 rows (frame index in batch, x1, y1, x2, y2) -- in the reference they come from the tracker-side
    crop rule applied to the detections (deep_sort.py:119-129); the benchmark supplies synthetic ones."""

    def __init__(self, yolo: YoloEngine, reid: Optional[ReidEngine], max_rois: int = 0):
        self.yolo, self.reid = yolo, reid
        dev = yolo.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        B, H, W = yolo.batch, yolo.h, yolo.w
        self.stage = [torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.stage_ready = [torch.cuda.Event() for _ in range(2)]
        self.stage_free = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.det_host = [torch.empty(B, yolo.max_det, 6, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.cnt_host = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.max_rois = max_rois
        if reid is not None:
            self.rois_dev = [torch.zeros(reid.capacity, 5, dtype=torch.int32, device=dev) for _ in range(2)]
            self.rois_host = [torch.zeros(reid.capacity, 5, dtype=torch.int32).pin_memory() for _ in range(2)]
            self.feat_host = [torch.empty(reid.capacity, 512, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.n_rois = [0, 0]
        self.submitted = 0
        self.collected = 0
        for e in self.stage_free:
            e.record(yolo.plan.stream)

    def submit(self, frames_pinned: torch.Tensor, rois: Optional[np.ndarray] = None, seg_sizes=None) -> None:
        assert self.submitted - self.collected < 2, "two batches are already in flight: call collect() first"
        k = self.submitted & 1
        ys = self.yolo.plan.stream
        # H2D on the copy stream as soon as the staging slot was consumed by the batch two steps back
        self.copy_stream.wait_event(self.stage_free[k])
        with torch.cuda.stream(self.copy_stream):
            self.stage[k].copy_(frames_pinned, non_blocking=True)
            n = 0
            if self.reid is not None and rois is not None:
                n = int(rois.shape[0])
                self.rois_host[k][:n] = torch.from_numpy(np.ascontiguousarray(rois, dtype=np.int32))
                self.rois_dev[k].copy_(self.rois_host[k], non_blocking=True)
            self.stage_ready[k].record(self.copy_stream)
        self.n_rois[k] = n
        # compute stream: staging -> the plan's static input (device-to-device), detect, ReID, results to pinned memory
        ys.wait_event(self.stage_ready[k])
        with torch.cuda.stream(ys):
            self.yolo.frames.copy_(self.stage[k], non_blocking=True)
        self.yolo.forward()
        with torch.cuda.stream(ys):
            self.det_host[k].copy_(self.yolo.det, non_blocking=True)
            self.cnt_host[k].copy_(self.yolo.det_count, non_blocking=True)
        if self.reid is not None and n > 0:
            rs = self.reid.stream
            rs.wait_stream(ys)
            with torch.cuda.stream(rs):
                self.reid.rois.copy_(self.rois_dev[k], non_blocking=True)
            self.reid.run(self.yolo.frames, None, n=n, seg_sizes=seg_sizes)
            with torch.cuda.stream(rs):
                self.feat_host[k][:n].copy_(self.reid.features[:n], non_blocking=True)
            ys.wait_stream(rs)
        self.stage_free[k].record(ys)
        self.done[k].record(ys)
        self.submitted += 1

    def collect(self) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
        assert self.collected < self.submitted, "nothing in flight"
        k = self.collected & 1
        self.done[k].synchronize()
        self.collected += 1
        n = self.n_rois[k]
        feats = self.feat_host[k][:n].numpy() if (self.reid is not None and n > 0) else None
        return self.det_host[k].numpy(), self.cnt_host[k].numpy(), feats
