"""Host frame source with the reference's surface (/root/reference/modules/datasets.py:14-94): `VideoLoader(config, path)` iterates a
video file through cv2.VideoCapture and yields the batch dicts the stages consume -- {'imgs': [RGB HWC uint8], 'frames': [1-based
frame ids], 'ori_imgs': [BGR HWC uint8]} -- one frame per batch as the reference's DataLoader(batch_size=1) does (`batch_size` is
an extension: the detector stage takes any number of frames per call).  Decoding stays on the host."""
from __future__ import annotations

import os
from typing import Dict, Iterator, Optional


class VideoSet:
    def __init__(self, config, input_path: str):
        self.input_path = input_path
        self.image_size = getattr(config, "image_size", None)
        self.initialize_stream()

    def initialize_stream(self) -> None:
        import cv2
        self.stream = cv2.VideoCapture(self.input_path)
        self.current_frame_id = 0
        if not self.stream.isOpened():
            raise AssertionError(f"Cannot read video {os.path.basename(self.input_path)}")
        self.WIDTH = int(self.stream.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.HEIGHT = int(self.stream.get(cv2.CAP_PROP_FRAME_HEIGHT))
        self.FPS = int(self.stream.get(cv2.CAP_PROP_FPS))
        self.NUM_FRAMES = int(self.stream.get(cv2.CAP_PROP_FRAME_COUNT))
        self.video_info = {"name": os.path.basename(self.input_path), "width": self.WIDTH, "height": self.HEIGHT, "fps": self.FPS,
                           "num_frames": self.NUM_FRAMES}

    def _pinned_slot(self):
        """$VCB_PINNED_FRAMES=N (N >= 2): frames are decoded into a ring of N page-locked (BGR, RGB) buffer pairs, so the stages copy
        them to the device in place (vcb_h2d_frames_inplace: no gather through a staging buffer).  A frame handed out stays valid
        for the next N - 1 reads: N must exceed the number of frames the consumer holds at once (batch size x batches in flight).
        Off by default: the reference hands out fresh arrays that stay valid for ever."""
        n = int(os.environ.get("VCB_PINNED_FRAMES", "0") or 0)
        if n < 2:
            return None
        ring = getattr(self, "_ring", None)
        if ring is None or ring[0].shape[0] != n:
            import torch
            if not torch.cuda.is_available():
                return None
            ring = self._ring = tuple(torch.empty(n, self.HEIGHT, self.WIDTH, 3, dtype=torch.uint8).pin_memory() for _ in range(2))
            self._ring_np = tuple(t.numpy() for t in ring)
        k = self.current_frame_id % n
        return self._ring_np[0][k], self._ring_np[1][k]

    def read(self) -> Optional[Dict]:
        import cv2
        slot = self._pinned_slot()
        if slot is None:
            ok, ori = self.stream.read()
        else:
            ok, got = self.stream.read(slot[0])
            ori = slot[0]
            if ok and got is not ori:                   # the decoder allocated its own array (size / type mismatch): copy once
                if got.shape != ori.shape:
                    slot = None
                    ori = got
                else:
                    ori[...] = got
        self.current_frame_id += 1
        if not ok:
            return None
        rgb = cv2.cvtColor(ori, cv2.COLOR_BGR2RGB) if slot is None else cv2.cvtColor(ori, cv2.COLOR_BGR2RGB, dst=slot[1])
        return {"img": rgb, "frame": self.current_frame_id, "ori_img": ori}

    def __len__(self) -> int:
        return self.NUM_FRAMES


class VideoLoader:
    def __init__(self, config, video_path: str, batch_size: int = 1):
        self.video_path = video_path
        self.dataset = VideoSet(config, video_path)
        self.video_info = self.dataset.video_info
        self.batch_size = batch_size

    def __len__(self) -> int:
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Optional[Dict]]:
        for _ in range(len(self)):
            items = [x for x in (self.dataset.read() for _ in range(self.batch_size)) if x is not None]
            if not items:
                yield None                      # datasets.py:68-70: a batch with no readable frame collates to None
                continue
            yield {"imgs": [s["img"] for s in items], "frames": [s["frame"] for s in items], "ori_imgs": [s["ori_img"] for s in items]}

    def reinitialize_stream(self) -> None:
        self.dataset.initialize_stream()
