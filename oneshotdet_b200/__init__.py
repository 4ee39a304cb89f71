"""oneshotdet_b200: B200-native (sm_100a) implementation of OneshotDet's inference hot path --
support->target matching on the FPN levels, FCOS post-processing and batched NMS -- behind the
reference's own interfaces (``_C.nms`` / ``layers.nms``, ``FCOSPostProcessor``, a matching module).

The compute lives in libosd_b200.so (C ABI, include/osd_b200.h); this package is the Python host side.
There is no CPU path and no eager-PyTorch fallback."""
from . import _C  # noqa: F401
from .layers import nms  # noqa: F401
from .modeling.matching import MatchingModule  # noqa: F401
from .modeling.rpn.fcos.inference import FCOSPostProcessor, make_fcos_postprocessor  # noqa: F401
from .modeling.poolers import LevelMapper, Pooler, make_pooler, roi_pool  # noqa: F401
from .modeling.roi_heads.box_head.box_head import BoxHeadDense  # noqa: F401
from .modeling.roi_heads.box_head.inference import BoxCoder, PostProcessor, make_roi_box_post_processor  # noqa: F401
from .modeling.support_pooling import SuppAlignLayer, SuppAvgPool, support_pool  # noqa: F401
from .ops import batched_nms, box_postprocess, fcos_postprocess, match_forward  # noqa: F401
from .structures.bounding_box import BoxList  # noqa: F401
from .structures.boxlist_ops import boxlist_nms, cat_boxlist, remove_small_boxes  # noqa: F401

__version__ = "0.1.0"
