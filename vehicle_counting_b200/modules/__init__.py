"""Drop-in for the hot-path stage wrappers of the reference's `modules` package."""
from .detect import ImageDetect
from .track import VideoTracker

__all__ = ["ImageDetect", "VideoTracker"]
