"""Stand-in for the reference's native module ``maskrcnn_benchmark._C`` (csrc/vision.cpp:7-15), restricted
to the operators that are on the hot path.  ``from oneshotdet_b200 import _C; _C.nms(dets, scores, thr)``
has the reference's signature and return contract (csrc/nms.h:10-28); ``_C.match_forward`` and ``_C.fcos_postprocess``
are the tensor-in / tensor-out forms of the matching module's forward and of ``FCOSPostProcessor.forward``.

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


def match_forward(features, supp_pooled, batch_size, mode="product"):
    """list of [B,C,H,W] (concat modes: [B,2C,H,W]); see csrc/torch_ext.cpp / ops.match_forward."""
    if _C_torch is not None:
        return _C_torch.match_forward(list(features), list(supp_pooled), int(batch_size), mode)
    from .ops import match_forward as _mf

    return _mf(features, supp_pooled, batch_size, mode)


def fcos_postprocess(box_cls, box_regression, centerness, image_sizes, strides, pre_nms_thresh, pre_nms_top_n, nms_thresh,
                     fpn_post_nms_top_n, min_size, strict=False, early_exit=True):
    """(boxes [B,K,4], scores [B,K], index int32 [B,K], count int32 [B]) -- FCOSPostProcessor.forward on tensors."""
    if _C_torch is not None:
        return _C_torch.fcos_postprocess(list(box_cls), list(box_regression), list(centerness),
                                         [(int(h), int(w)) for h, w in image_sizes], [int(s) for s in strides],
                                         float(pre_nms_thresh), int(pre_nms_top_n), float(nms_thresh), int(fpn_post_nms_top_n),
                                         float(min_size), bool(strict), bool(early_exit))
    from .ops import fcos_postprocess as _fp

    r = _fp(box_cls, box_regression, centerness, strides, image_sizes, pre_nms_thresh, pre_nms_top_n, nms_thresh,
            fpn_post_nms_top_n, min_size, strict, early_exit)
    return r.boxes, r.scores, r.index, r.count


__all__ = ["nms", "match_forward", "fcos_postprocess", "BINDING"]
