"""Mirror of ``maskrcnn_benchmark.layers`` for the hot path (layers/__init__.py:9 re-exports nms)."""
from .nms import nms  # noqa: F401

__all__ = ["nms"]
