"""Support-embedding producers, the step right before the matching path.

``SuppAlignLayer`` mirrors maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:20-52 (same constructor
arguments, same ``forward(x, boxes)``): per FPN level a ROIAlign with a (1,1) output over one box per support image.
``SuppAvgPool`` stands in for ``nn.AdaptiveAvgPool2d((1,1))`` applied per level (generalized_rcnn.py:94, :303).
Both return ``[N, C, 1, 1]`` tensors per level -- exactly what ``MatchingModule.forward`` takes as ``supp_pooled``."""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from oneshotdet_b200 import _lib
from oneshotdet_b200._lib import OsdError, SupportPoolDesc, OSD_MAX_LEVELS

POOL_MODES = {"roialign": 0, "avg": 1}


@torch.no_grad()
def support_pool(features, rois=None, scales=None, sampling_ratio: int = 2, mode: str = "roialign"):
    """features[l] [N,C,H_l,W_l] fp32 CUDA (NCHW); rois [N,4] (x1,y1,x2,y2) in support-image coordinates and
    scales[l] for mode 'roialign'.  One kernel launch for all levels.  Returns a list of [N,C,1,1]."""
    if mode not in POOL_MODES:
        raise OsdError(f"support_pool: unknown mode '{mode}'")
    lib = _lib.load()
    nl = len(features)
    if nl == 0 or nl > OSD_MAX_LEVELS:
        raise OsdError("support_pool: need 1..8 levels")
    dev = features[0].device
    _lib.require_device(dev)
    n, c = features[0].shape[:2]
    d = SupportPoolDesc()
    d.num_levels, d.num_supports, d.channels = nl, n, c
    d.mode, d.sampling_ratio = POOL_MODES[mode], int(sampling_ratio)
    keep, outs = [], []
    if mode == "roialign":
        if rois is None or scales is None or len(scales) != nl:
            raise OsdError("support_pool: mode 'roialign' needs rois [N,4] and one scale per level")
        rois = rois.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(rois.shape) != (n, 4):
            raise OsdError(f"support_pool: rois must be [{n},4], got {tuple(rois.shape)}")
        d.rois = rois.data_ptr()
        keep.append(rois)
    for l, f in enumerate(features):
        if f.dim() != 4 or f.size(0) != n or f.size(1) != c or f.dtype != torch.float32 or f.device != dev:
            raise OsdError(f"support_pool: level {l}: features must be [N={n},C={c},H,W] float32 on {dev}")
        f = f.contiguous()
        o = torch.empty((n, c, 1, 1), dtype=torch.float32, device=dev)
        d.height[l], d.width[l] = f.shape[2], f.shape[3]
        d.spatial_scale[l] = float(scales[l]) if scales is not None else 1.0
        d.feat[l], d.out[l] = f.data_ptr(), o.data_ptr()
        keep.append(f)
        outs.append(o)
    with torch.cuda.device(dev):
        rc = lib.osd_support_pool(ctypes.byref(d), _lib.current_stream_ptr(dev))
    _lib.check(rc, "osd_support_pool")
    return outs


class SuppAlignLayer(nn.Module):
    def __init__(self, scales, output_size, sampling_ratio):
        super().__init__()
        if tuple(output_size) != (1, 1):
            raise NotImplementedError("SuppAlignLayer on the accelerated path pools to (1,1) (generalized_rcnn.py:88-92)")
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = int(sampling_ratio)

    def convert_to_roi_format(self, boxes):
        """One box per support image, image-major (generalized_rcnn.py:32-45)."""
        if any(len(b.bbox) != 1 for b in boxes):
            raise NotImplementedError("exactly one box per support image is supported (the whole-image box of "
                                      "generalized_rcnn.py:257)")
        return torch.cat([b.bbox for b in boxes], dim=0)

    def forward(self, x, boxes):
        rois = self.convert_to_roi_format(boxes)
        return support_pool(list(x), rois, self.scales[:len(x)], self.sampling_ratio, "roialign")


class SuppAvgPool(nn.Module):
    """AdaptiveAvgPool2d((1,1)) over every level in one launch (the SUPP_ROIALIGN=False branch)."""

    def forward(self, x):
        return support_pool(list(x), mode="avg")
