"""`DeepSort` / `Extractor` with the reference's signatures (/root/reference/networks/deepsort/deep_sort.py:14-129,
deep/feature_extractor.py:9-47).  The appearance path (crop -> resize -> normalise -> CNN -> L2 norm) runs on
the GPU through `ReidEngine`; the association step is sort.py (host, as in the reference).

bn_mode: "train" reproduces the reference as shipped (its Extractor never calls .eval(), so BatchNorm uses the
statistics of the crops of ONE call = all detections of one class in one frame); "eval" is the folded-BN fast
path.  Default from $VCB_REID_BN (default "train", i.e. reference-faithful)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ...engine import ReidEngine
from ...weights import load_reid_state_dict, synth_reid_state_dict
from .sort import Detection, Gallery, Tracker, host_nms

__all__ = ["DeepSort", "Extractor"]

_ENGINES: Dict[Tuple[str, str, int], ReidEngine] = {}      # (weights path, bn_mode, device) -> shared engine


def _shared_engine(model_path: str, bn_mode: str, capacity: Optional[int] = None) -> ReidEngine:
    """The reference builds one Extractor (one copy of the net) per class tracker (modules/track.py:16 ->
    80 copies for COCO); the weights are identical, so all instances share one engine."""
    dev = torch.cuda.current_device()
    key = (model_path, bn_mode, dev)
    if capacity is None:
        capacity = int(os.environ.get("VCB_REID_CAPACITY", "512"))       # crops per engine pass (batched callers raise it)
    if key not in _ENGINES:
        sd = synth_reid_state_dict(0) if model_path.startswith("synthetic") else load_reid_state_dict(model_path)
        _ENGINES[key] = ReidEngine(sd, capacity=capacity, device=f"cuda:{dev}", bn_mode=bn_mode)
    return _ENGINES[key]


class Extractor:
    """feature_extractor.py:9-47: `Extractor(model_path, use_cuda)(im_crops) -> float32 [n, 512]`."""

    def __init__(self, model_path, use_cuda=True, bn_mode: Optional[str] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("the ReID path runs on the GPU only (no CPU fallback)")
        self.bn_mode = bn_mode or os.environ.get("VCB_REID_BN", "train")
        self.engine = _shared_engine(model_path, self.bn_mode)
        self.size = (50, 50)

    def __call__(self, im_crops: Sequence[np.ndarray]) -> np.ndarray:
        n = len(im_crops)
        if n == 0:
            return np.zeros((0, 512), np.float32)
        # the crops are stacked into ONE atlas "frame" (rows of crop i at [y_i, y_i + h_i), columns [0, w_i)) inside an engine-owned
        # staging buffer that only ever grows; its height / width change from call to call, so this path launches eagerly instead
        # of replaying a graph (nothing is captured, packed or allocated per call)
        wm = max(c.shape[1] for c in im_crops)
        rois = np.zeros((n, 5), np.int32)
        y = 0
        for i, c in enumerate(im_crops):
            if c.shape[0] == 0 or c.shape[1] == 0:
                raise ValueError("empty crop (the reference fails inside cv2.resize here)")
            rois[i] = (0, 0, y, c.shape[1], y + c.shape[0])
            y += c.shape[0]
        eng = self.engine
        host, dev = eng.atlas(y * wm * 3)
        eng.stream.synchronize()                        # the previous upload out of the pinned atlas has been consumed
        av = host[:y * wm * 3].numpy().reshape(y, wm, 3)
        for i, c in enumerate(im_crops):
            av[rois[i, 2]:rois[i, 4], :c.shape[1]] = c
        with torch.cuda.stream(eng.stream):
            dev[:y * wm * 3].copy_(host[:y * wm * 3], non_blocking=True)
        eng.run(dev[:y * wm * 3].view(1, y, wm, 3), rois, seg_sizes=[n], use_graph=False)
        return eng.download(n)

    def from_frame(self, frame: np.ndarray, rois_xyxy: np.ndarray, frame_dev: Optional[torch.Tensor] = None) -> np.ndarray:
        """Same result as __call__ on `frame[y1:y2, x1:x2]` crops, without materialising them on the host.  `frame_dev`: the frame
        already on the device (VideoTracker uploads it once per frame and hands it to every class tracker); otherwise it is
        uploaded here, every call -- there is no content-based guess about whether this is "the same" frame as last time."""
        n = len(rois_xyxy)
        dev = frame_dev if frame_dev is not None else self.engine.stage_frames(frame)
        rois = np.concatenate([np.zeros((n, 1), np.int32), np.asarray(rois_xyxy, np.int32)], 1)
        self.engine.run(dev, rois, seg_sizes=[n])
        return self.engine.download(n)


    def from_frames(self, frames: Sequence[np.ndarray], boxes_xyxy: Sequence[np.ndarray]) -> List[np.ndarray]:
        """Batched form of the per-frame call chain DeepSort.update -> _get_features -> Extractor (deep_sort.py:25-31, :119-129)
        for a list of equally sized BGR frames: `boxes_xyxy[i]` are float64 xyxy boxes of frame i, the crop rectangles follow the
        reference rule, all crops go through ONE engine pass (train-mode BatchNorm: one statistics segment per frame = one
        reference call per frame) and the embeddings come back as one float32 [n_i, 512] array per frame."""
        h, w = frames[0].shape[:2]
        seg = [len(b) for b in boxes_xyxy]
        b = np.concatenate([np.asarray(x, np.float64).reshape(-1, 4) for x in boxes_xyxy], 0)
        bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
        cx, cy = b[:, 0] + bw / 2, b[:, 1] + bh / 2
        # deep_sort.py:89-95: int() truncation (towards zero) of centre -/+ half size, clipped to the frame
        x1 = np.maximum(np.trunc(cx - bw / 2), 0); x2 = np.minimum(np.trunc(cx + bw / 2), w - 1)
        y1 = np.maximum(np.trunc(cy - bh / 2), 0); y2 = np.minimum(np.trunc(cy + bh / 2), h - 1)
        if ((x2 <= x1) | (y2 <= y1)).any():
            raise ValueError("empty crop (the reference fails inside cv2.resize here)")
        rois = np.stack([np.repeat(np.arange(len(seg)), seg), x1, y1, x2, y2], 1).astype(np.int32)
        n = len(rois)
        eng = self.engine
        if n > eng.capacity:
            raise ValueError(f"{n} crops in one call; the shared ReID engine holds {eng.capacity}")
        dev = eng.stage_frame_list(frames)
        seg_nz = [k for k in seg if k > 0]
        eng.run(dev, rois, seg_sizes=seg_nz)
        feats = eng.download(n)
        out, off = [], 0
        for k in seg:
            out.append(feats[off:off + k])
            off += k
        return out


class DeepSort:
    def __init__(self, model_path, max_dist=0.2, min_confidence=0.3, nms_max_overlap=1.0, max_iou_distance=0.7, max_age=70,
                 n_init=3, nn_budget=100, use_cuda=True, bn_mode: Optional[str] = None):
        self.min_confidence = min_confidence
        self.nms_max_overlap = nms_max_overlap
        self.extractor = Extractor(model_path, use_cuda=use_cuda, bn_mode=bn_mode)
        self.tracker = Tracker(Gallery(max_dist, nn_budget), max_iou_distance=max_iou_distance, max_age=max_age, n_init=n_init)

    # -- box conventions (deep_sort.py:67-117) ---------------------------------------------------
    @staticmethod
    def _xyxy_to_xywh(bbox_xyxy: np.ndarray) -> np.ndarray:
        b = np.array(bbox_xyxy, dtype=np.float64, copy=True)
        b[:, 2] = bbox_xyxy[:, 2] - bbox_xyxy[:, 0]
        b[:, 3] = bbox_xyxy[:, 3] - bbox_xyxy[:, 1]
        b[:, 0] = b[:, 0] + b[:, 2] / 2
        b[:, 1] = b[:, 1] + b[:, 3] / 2
        return b

    def _crop_rect(self, box_xywh) -> Tuple[int, int, int, int]:
        x, y, w, h = box_xywh
        x1 = max(int(x - w / 2), 0)
        x2 = min(int(x + w / 2), self.width - 1)
        y1 = max(int(y - h / 2), 0)
        y2 = min(int(y + h / 2), self.height - 1)
        return x1, y1, x2, y2

    def _tlwh_to_xyxy(self, tlwh) -> Tuple[int, int, int, int]:
        x, y, w, h = tlwh
        return max(int(x), 0), max(int(y), 0), min(int(x + w), self.width - 1), min(int(y + h), self.height - 1)

    def _get_features(self, bbox_xywh: np.ndarray, ori_img: np.ndarray, frame_dev: Optional[torch.Tensor] = None) -> np.ndarray:
        """deep_sort.py:119-129: crops come from the BGR original-resolution frame."""
        rects = np.array([self._crop_rect(b) for b in bbox_xywh], np.int32).reshape(-1, 4)
        if len(rects) == 0:
            return np.array([])
        if ((rects[:, 2] <= rects[:, 0]) | (rects[:, 3] <= rects[:, 1])).any():
            raise ValueError("empty crop (the reference fails inside cv2.resize here)")
        return self.extractor.from_frame(ori_img, rects, frame_dev)

    def update(self, bbox_xyxy, confidences, ori_img, features: Optional[np.ndarray] = None, frame_dev: Optional[torch.Tensor] = None):
        """deep_sort.py:25-59.  `features` (optional, not in the reference) injects precomputed embeddings; `frame_dev` (optional)
        is `ori_img` already uploaded by the caller."""
        self.height, self.width = ori_img.shape[:2]
        bbox_xyxy = np.asarray(bbox_xyxy, dtype=np.float64)
        bbox_xywh = self._xyxy_to_xywh(bbox_xyxy)
        if features is None:
            features = self._get_features(bbox_xywh, ori_img, frame_dev)
        tlwh = bbox_xywh.copy()
        tlwh[:, 0] = bbox_xywh[:, 0] - bbox_xywh[:, 2] / 2.0
        tlwh[:, 1] = bbox_xywh[:, 1] - bbox_xywh[:, 3] / 2.0
        dets = [Detection(tlwh[i], c, features[i]) for i, c in enumerate(confidences) if c > self.min_confidence]
        boxes = np.array([d.tlwh for d in dets])
        scores = np.array([d.confidence for d in dets])
        dets = [dets[i] for i in host_nms(boxes, self.nms_max_overlap, scores)]
        self.tracker.predict()
        self.tracker.update(dets)
        outputs = []
        for t in self.tracker.tracks:
            if not t.is_confirmed() or t.time_since_update > 1:
                continue
            x1, y1, x2, y2 = self._tlwh_to_xyxy(t.to_tlwh())
            # column 5 is always -1 (features were moved to the gallery), column 6 = int(score) (deep_sort.py:52-56)
            outputs.append(np.array([x1, y1, x2, y2, t.track_id, -1, int(t.get_confidence_score())],
                                    dtype=np.int64))
        if len(outputs) > 0:
            outputs = np.stack(outputs, axis=0)
        return outputs
