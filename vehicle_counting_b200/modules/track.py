"""`VideoTracker` with the reference's surface (/root/reference/modules/track.py:8-70): one `DeepSort` per class, per-frame
fan-out of the detections by label, rows collected into {'tracks', 'boxes', 'labels', 'scores'}.

The reference calls each class tracker's Extractor separately (one H2D, one tiny-batch CNN pass and one D2H per class with
detections, track.py:50-59).  Here the crops of ALL classes of the frame go through the ReID engine in ONE pass: the frame
is uploaded once, the crop rectangles follow the reference rule (deep_sort.py:78-95, :119-125), and with train-mode
BatchNorm each class is one statistics segment, which is exactly the reference's per-call batch, so the
embeddings -- and therefore track ids and boxes -- are those of the per-class calls.  `VideoCounting` (zone filter, movement
direction, CSV; track.py:72-137) is once-per-video host post-processing, restated in ../counting.py."""
from __future__ import annotations

import random

import numpy as np

from ..counting import PALETTE, check_bbox_intersect_polygon, find_best_match_direction, load_zone_anno, save_tracking_to_csv
from ..networks import DeepSort

__all__ = ["VideoTracker", "VideoCounting"]


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

    def _features_all_classes(self, frame_dev, image, per_class):
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
        rois = np.concatenate([np.zeros((n, 1), np.int32), rects], 1)
        eng.run(frame_dev, rois, seg_sizes=seg)
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
        feats, frame_dev = {}, None
        if per_class:
            # the frame goes up ONCE per run() into the shared engine's staging buffer; every class tracker gets the device copy
            frame_dev = self.deepsort[per_class[0][0]].extractor.engine.stage_frames(image)
            feats = self._features_all_classes(frame_dev, image, [(i, b) for i, b, _ in per_class])
        for i, b, s in per_class:
            # output rows: x1, y1, x2, y2, track_id, -1, int(score)
            outputs = self.deepsort[i].update(b, s, image, features=None if feats is None else feats[i], frame_dev=frame_dev)
            for obj in outputs:
                result_dict["tracks"].append(obj[4])
                result_dict["boxes"].append(obj[:4])
                result_dict["labels"].append(i)
        result_dict["boxes"] = np.array(result_dict["boxes"])
        return result_dict


class VideoCounting:
    """track.py:72-137: keeps, per label and track, the boxes (xyxy) whose corners touch the zone polygon, assigns each track the
    annotated direction closest (cosine) to its first-centre -> last-centre vector and writes the tracking CSV."""

    def __init__(self, class_names, zone_path, minimum_length=4) -> None:
        self.class_names = class_names
        self.num_classes = len(class_names)
        self.track_dict = [{} for _ in range(self.num_classes)]
        self.minimum_length = minimum_length
        self.zone_path = zone_path
        self.polygons, self.directions = load_zone_anno(zone_path)

    def run(self, frames, tracks, labels, boxes, output_path=None):
        for frame_id, track_id, label_id, box in zip(frames, tracks, labels, boxes):
            if not check_bbox_intersect_polygon(self.polygons, box):
                continue
            per_label = self.track_dict[label_id]
            if track_id not in per_label:
                per_label[track_id] = {"boxes": [], "frames": [], "color": random.sample(PALETTE, 1)[0]}
            per_label[track_id]["boxes"].append(box)
            per_label[track_id]["frames"].append(frame_id)
        for per_label in self.track_dict:
            for rec in per_label.values():
                b0, b1 = rec["boxes"][0], rec["boxes"][-1]
                first = ((b0[2] + b0[0]) / 2, (b0[3] + b0[1]) / 2)
                last = ((b1[2] + b1[0]) / 2, (b1[3] + b1[1]) / 2)
                rec["direction"] = find_best_match_direction(obj_vector=(first, last), paths=self.directions)
        if output_path is not None:
            save_tracking_to_csv(self.track_dict, output_path)
        return self.track_dict                                  # track.py:137; CountingPipeline keeps it as `result_dict`
