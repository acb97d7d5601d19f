"""Drop-in for the hot-path stage wrapper of the reference's `modules` package."""
from .detect import ImageDetect

__all__ = ["ImageDetect"]
