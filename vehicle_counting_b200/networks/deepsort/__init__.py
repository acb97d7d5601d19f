from .deep_sort import DeepSort, Extractor

__all__ = ["DeepSort", "Extractor"]
