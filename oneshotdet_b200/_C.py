"""Stand-in for the reference's native module ``maskrcnn_benchmark._C`` (csrc/vision.cpp:7-15), restricted
to the operator that is on the hot path.  ``from oneshotdet_b200 import _C; _C.nms(dets, scores, thr)``
has the reference's signature and return contract (csrc/nms.h:10-28).

Two bindings reach the same kernels in libosd_b200.so: the thin torch C++ extension ``_C_torch``
(csrc/torch_ext.cpp, built by ``python -m oneshotdet_b200.build --ext`` / ``__graft_entry__.build()``) and the
ctypes binding in ``ops.py``.  ``nms`` is the extension's when it has been built, else the ctypes one; neither
has a CPU path."""
from .ops import nms as _nms_ctypes

try:
    from . import _C_torch  # type: ignore[attr-defined]
except ImportError:  # extension not built: same kernels through ctypes
    _C_torch = None

nms = _C_torch.nms if _C_torch is not None else _nms_ctypes
BINDING = "torch-extension" if _C_torch is not None else "ctypes"

__all__ = ["nms", "BINDING"]
