"""Stand-in for the reference's native module ``maskrcnn_benchmark._C`` (csrc/vision.cpp:7-15), restricted
to the operator that is on the hot path.  ``from oneshotdet_b200 import _C; _C.nms(dets, scores, thr)``
has the reference's signature and return contract (csrc/nms.h:10-28)."""
from .ops import nms  # noqa: F401

__all__ = ["nms"]
