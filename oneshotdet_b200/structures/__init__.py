from .bounding_box import BoxList  # noqa: F401
from .boxlist_ops import boxlist_nms, cat_boxlist, remove_small_boxes  # noqa: F401
