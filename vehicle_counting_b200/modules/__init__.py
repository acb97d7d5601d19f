"""Drop-in for the stage wrappers of the reference's `modules` package (detect.py, track.py)."""
from .detect import ImageDetect
from .track import VideoCounting, VideoTracker

__all__ = ["ImageDetect", "VideoTracker", "VideoCounting"]
