"""`VideoTracker` with the reference's surface (/root/reference/modules/track.py:8-70): one `DeepSort` per class, per-frame
fan-out of the detections by label, rows collected into {'tracks', 'boxes', 'labels', 'scores'}.

The reference calls each class tracker's Extractor separately (one H2D, one tiny-batch CNN pass and one D2H per class with
detections, track.py:50-59).  Here the crops of ALL classes of the frame go through the ReID engine in ONE pass: the frame
is uploaded once, the crop rectangles follow the reference rule (deep_sort.py:78-95, :119-125), and with train-mode
BatchNorm each class is one statistics segment, which is exactly the reference's per-call batch, so the
embeddings -- and therefore track ids and boxes -- are those of the per-class calls.  `VideoCounting` (zone filter, CSV) is
host post-processing outside the hot path and is not mirrored."""
from __future__ import annotations

import numpy as np

from ..networks import DeepSort
from ..networks.deepsort.deep_sort import _FRAMES

__all__ = ["VideoTracker"]


class VideoTracker:
    def __init__(self, num_classes, cam_config, video_info, deepsort_chepoint, bn_mode=None):
        tracking_config = cam_config["tracking_config"]
        self.num_classes = num_classes
        self.video_info = video_info
        self.num_frames = video_info["num_frames"]
        self.bn_mode = bn_mode
        # a tracker (Kalman state, gallery, id counter) per class, as in the reference; the ReID weights are shared
        self.deepsort = [self.build_tracker(deepsort_chepoint, tracking_config) for _ in range(num_classes)]

    def build_tracker(self, checkpoint, cam_cfg):
        return DeepSort(checkpoint, max_dist=cam_cfg["MAX_DIST"], min_confidence=cam_cfg["MIN_CONFIDENCE"],
                        nms_max_overlap=cam_cfg["NMS_MAX_OVERLAP"], max_iou_distance=cam_cfg["MAX_IOU_DISTANCE"],
                        max_age=cam_cfg["MAX_AGE"], n_init=cam_cfg["N_INIT"], nn_budget=cam_cfg["NN_BUDGET"], use_cuda=1,
                        bn_mode=self.bn_mode)

    def _features_all_classes(self, image, per_class):
        """one ReID pass for the whole frame; per_class: list of (class id, xyxy float64 [n,4]) -> {class id: float32 [n,512]}"""
        ds0 = self.deepsort[per_class[0][0]]
        eng = ds0.extractor.engine
        rects, seg = [], []
        for i, xyxy in per_class:
            ds = self.deepsort[i]
            ds.height, ds.width = image.shape[:2]
            r = np.array([ds._crop_rect(b) for b in ds._xyxy_to_xywh(xyxy)], np.int32).reshape(-1, 4)
            if ((r[:, 2] <= r[:, 0]) | (r[:, 3] <= r[:, 1])).any():
                raise ValueError("empty crop (the reference fails inside cv2.resize here)")
            rects.append(r)
            seg.append(len(r))
        rects = np.concatenate(rects, 0)
        n = len(rects)
        if n > eng.capacity:        # more detections than the shared engine was sized for: fall back to per-class calls
            return None
        dev = _FRAMES.get(image, eng.stream)
        rois = np.concatenate([np.zeros((n, 1), np.int32), rects], 1)
        eng.run(dev, rois, seg_sizes=seg)
        feats = eng.download(n)
        out, off = {}, 0
        for (i, _), k in zip(per_class, seg):
            out[i] = feats[off:off + k]
            off += k
        return out

    def run(self, image, boxes, labels, scores):
        """track.py:30-70: boxes are xywh (top-left) in original pixels, image is the BGR frame."""
        self.obj_track = [{} for _ in range(self.num_classes)]
        bbox_xyxy = np.array(boxes, dtype=np.float64, copy=True)
        bbox_xyxy[:, 2] += bbox_xyxy[:, 0]
        bbox_xyxy[:, 3] += bbox_xyxy[:, 1]
        result_dict = {"tracks": [], "boxes": [], "labels": [], "scores": []}
        labels = np.asarray(labels)
        scores = np.asarray(scores)
        per_class = []
        for i in range(self.num_classes):
            mask = labels == i
            if mask.any():
                per_class.append((i, bbox_xyxy[mask], scores[mask]))
        feats = self._features_all_classes(image, [(i, b) for i, b, _ in per_class]) if per_class else {}
        for i, b, s in per_class:
            # output rows: x1, y1, x2, y2, track_id, -1, int(score)
            outputs = self.deepsort[i].update(b, s, image, features=None if feats is None else feats[i])
            for obj in outputs:
                result_dict["tracks"].append(obj[4])
                result_dict["boxes"].append(obj[:4])
                result_dict["labels"].append(i)
        result_dict["boxes"] = np.array(result_dict["boxes"])
        return result_dict
