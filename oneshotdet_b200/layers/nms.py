"""Mirror of maskrcnn_benchmark/layers/nms.py:3-5: ``nms = _C.nms``."""
from oneshotdet_b200 import _C

nms = _C.nms
