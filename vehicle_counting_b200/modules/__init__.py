"""Drop-in for the stage wrappers of the reference's `modules` package (detect.py, track.py) and for its per-video driver
(/root/reference/modules/__init__.py:8-100)."""
from __future__ import annotations

import os

from .datasets import VideoLoader
from .detect import ImageDetect
from .track import VideoCounting, VideoTracker

__all__ = ["ImageDetect", "VideoTracker", "VideoCounting", "VideoLoader", "CountingPipeline"]


class CountingPipeline:
    """modules/__init__.py:8-100 with the same constructor and the same per-frame loop: detect -> (frames with detections) track ->
    collect rows -> VideoCounting -> `<output_path>/<cam_name>.csv`.  The overlay video the reference renders afterwards
    (VideoWriter.write_full_to_video, :95-100) is presentation, outside the hot path, and is not produced."""

    def __init__(self, args, config, cam_config, batch_size: int = 1):
        self.detector = ImageDetect(args, config)
        self.class_names = self.detector.class_names
        self.video_path = args.input_path
        self.saved_path = args.output_path
        self.cam_config = cam_config
        self.zone_path = cam_config.zone_path
        self.config = config
        self.batch_size = batch_size
        if os.path.isdir(self.video_path):
            self.all_video_paths = [os.path.join(self.video_path, i) for i in sorted(os.listdir(self.video_path))]
        else:
            self.all_video_paths = [self.video_path]

    def get_cam_name(self, path):
        return os.path.basename(path)[:-4]

    def run(self):
        results = {}
        for video_path in self.all_video_paths:
            cam_name = self.get_cam_name(video_path)
            videoloader = VideoLoader(self.config, video_path, batch_size=self.batch_size)
            self.tracker = VideoTracker(len(self.class_names), self.cam_config.cam[cam_name], videoloader.dataset.video_info,
                                        deepsort_chepoint=self.cam_config.checkpoint)
            videocounter = VideoCounting(class_names=self.class_names, zone_path=os.path.join(self.zone_path, cam_name + ".json"))
            obj_dict = {"frames": [], "tracks": [], "labels": [], "boxes": []}
            for batch in videoloader:
                if batch is None:
                    continue
                preds = self.detector.run(batch)
                ori_imgs = batch["ori_imgs"]
                for i in range(len(ori_imgs)):
                    boxes, labels, scores = preds["boxes"][i], preds["labels"][i], preds["scores"][i]
                    if len(boxes) == 0:
                        continue
                    track_result = self.tracker.run(ori_imgs[i], boxes, labels, scores)
                    for j in range(len(track_result["boxes"])):
                        obj_dict["frames"].append(batch["frames"][i])
                        obj_dict["tracks"].append(track_result["tracks"][j])
                        obj_dict["labels"].append(track_result["labels"][j])
                        obj_dict["boxes"].append(track_result["boxes"][j])
            os.makedirs(self.saved_path, exist_ok=True)
            results[cam_name] = videocounter.run(frames=obj_dict["frames"], tracks=obj_dict["tracks"], labels=obj_dict["labels"],
                                                 boxes=obj_dict["boxes"], output_path=os.path.join(self.saved_path, cam_name + ".csv"))
        return results
