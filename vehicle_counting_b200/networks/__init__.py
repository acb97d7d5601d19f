"""Drop-in for the reference's `networks` package (/root/reference/networks/__init__.py:1-3)."""
from .yolo import get_model, YoloBackbone
from .detector import Detector
from .deepsort import DeepSort

__all__ = ["get_model", "YoloBackbone", "Detector", "DeepSort"]
