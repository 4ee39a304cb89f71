"""Drop-in for the second stage's multi-level ROI pooler (maskrcnn_benchmark/modeling/poolers.py): ``LevelMapper``
(:10-41) and ``Pooler`` (:44-125) with the reference's constructor arguments and ``forward(x, boxes)`` contract, as ONE
kernel launch for all levels and ROIs (``osd_roi_pool``) instead of per-level nonzero / gather / ROIAlign / scatter.
SURVEY section 8(f) row 2 (pooling half).  fp32 results are bit-identical to the reference's CPU ROIAlign."""
from __future__ import annotations

import ctypes
import math

import torch
from torch import nn

from oneshotdet_b200 import _lib
from oneshotdet_b200._lib import OsdError, RoiPoolDesc, OSD_MAX_LEVELS


_workspace = _lib.Workspace()


@torch.no_grad()
def roi_pool(features, rois, scales, output_size: int = 7, sampling_ratio: int = 2, roi_count=None,
             canonical_scale: float = 224.0, canonical_level: int = 4, eps: float = 1e-6, return_levels: bool = False,
             channels_last: bool = True, rows_bf16: bool = False):
    """features[l] [B,C,H_l,W_l] fp32 CUDA (NCHW); rois [B,R,4] xyxy (image coordinates; ROI (b, r) reads image b);
    scales[l] per level.  Returns [B*R, C, P, P] (and the int32 level of every ROI when ``return_levels``).
    ``channels_last`` (default, needs C % 4 == 0): the library first transposes the maps to [B, H*W, C] in a scratch
    buffer so that every bilinear tap is one contiguous channel vector; False pools straight from NCHW.  Same bits.
    ``rows_bf16``: return bf16 [B*R, P*P, C] instead -- the same values rounded once to bf16 in the K-major row layout
    the dense head's GEMMs read (``BoxHeadDense``), skipping the fp32 [B*R,C,P,P] tensor altogether."""
    lib = _lib.load()
    nl = len(features)
    if nl == 0 or nl > OSD_MAX_LEVELS or len(scales) != nl:
        raise OsdError("roi_pool: need 1..8 levels and one scale per level")
    dev = features[0].device
    _lib.require_device(dev)
    if rois.dim() != 3 or rois.size(2) != 4 or rois.dtype != torch.float32 or rois.device != dev:
        raise OsdError(f"roi_pool: rois must be float32 [B,R,4] on {dev}, got {tuple(rois.shape)} {rois.dtype}")
    b, c = features[0].shape[:2]
    if rois.size(0) != b:
        raise OsdError(f"roi_pool: {b} images but rois for {rois.size(0)}")
    r = rois.size(1)
    d = RoiPoolDesc()
    d.num_levels, d.batch, d.rois_per_image, d.channels = nl, b, r, c
    d.pooled_size, d.sampling_ratio = int(output_size), int(sampling_ratio)
    # poolers.py:72-74: the level range follows from the first and last scale
    d.k_min = int(round(-math.log2(float(scales[0]))))
    d.k_max = int(round(-math.log2(float(scales[-1]))))
    d.canonical_scale, d.canonical_level, d.eps = float(canonical_scale), int(canonical_level), float(eps)
    keep = []
    for l, f in enumerate(features):
        if f.dim() != 4 or f.size(0) != b or f.size(1) != c or f.dtype != torch.float32 or f.device != dev:
            raise OsdError(f"roi_pool: level {l}: features must be [B={b},C={c},H,W] float32 on {dev}")
        f = f.contiguous()
        keep.append(f)
        d.height[l], d.width[l], d.spatial_scale[l], d.feat[l] = f.shape[2], f.shape[3], float(scales[l]), f.data_ptr()
    rois = rois.contiguous()
    if rows_bf16:
        if not (channels_last and c % 4 == 0):
            raise OsdError("roi_pool: rows_bf16 needs the channels-last path (C % 4 == 0)")
        out = torch.empty((b * r, d.pooled_size * d.pooled_size, c), dtype=torch.bfloat16, device=dev)
    else:
        out = torch.empty((b * r, c, d.pooled_size, d.pooled_size), dtype=torch.float32, device=dev)
    levels = torch.empty((b * r,), dtype=torch.int32, device=dev) if return_levels else None
    if roi_count is not None:
        if roi_count.dtype != torch.int32 or roi_count.device != dev or roi_count.numel() != b:
            raise OsdError("roi_pool: roi_count must be an int32 vector [B] on the inputs' device")
        roi_count = roi_count.contiguous()
        d.roi_count = roi_count.data_ptr()
    d.rois = rois.data_ptr()
    if rows_bf16:
        d.out, d.out_nhwc_bf16 = None, out.data_ptr()
    else:
        d.out = out.data_ptr()
    d.levels_out = levels.data_ptr() if levels is not None else None
    if channels_last and c % 4 == 0 and b * r > 0:  # pool from a channels-last copy of the maps made by the library
        need = ctypes.c_size_t(0)
        _lib.check(lib.osd_roi_pool_workspace_bytes(ctypes.byref(d), ctypes.byref(need)), "osd_roi_pool_workspace_bytes")
        ws = _workspace.get(dev, need.value)
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    if b * r > 0:
        with torch.cuda.device(dev):
            rc = lib.osd_roi_pool(ctypes.byref(d), _lib.current_stream_ptr(dev))
        _lib.check(rc, "osd_roi_pool")
    return (out, levels) if return_levels else out


class LevelMapper:
    """poolers.py:10-41, as a holder of the heuristic's constants: the mapping itself is evaluated inside the pooling
    kernel (``roi_pool(..., return_levels=True)`` exposes it)."""

    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min, self.k_max = k_min, k_max
        self.s0, self.lvl0, self.eps = canonical_scale, canonical_level, eps


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        if output_size[0] != output_size[1]:
            raise NotImplementedError("square pooled outputs only (POOLER_RESOLUTION is a single integer, poolers.py:128-135)")
        self.output_size = tuple(output_size)
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = int(sampling_ratio)
        lvl_min = -math.log2(self.scales[0])
        lvl_max = -math.log2(self.scales[-1])
        self.map_levels = LevelMapper(lvl_min, lvl_max)

    def forward_fixed(self, x, rois, roi_count=None, rows_bf16: bool = False):
        """rois [B,R,4] device tensor (e.g. the FCOS stage's padded output and its counts) -> [B,R,C,P,P]; no host sync.
        ``rows_bf16``: bf16 [B,R,P*P,C] for ``BoxHeadDense`` instead (see ``roi_pool``)."""
        m = self.map_levels
        out = roi_pool(list(x), rois, self.scales[:len(x)], self.output_size[0], self.sampling_ratio, roi_count,
                       m.s0, m.lvl0, m.eps, rows_bf16=rows_bf16)
        return out.view(rois.size(0), rois.size(1), out.size(1), out.size(2), *out.shape[3:])

    def forward(self, x, boxes):
        """poolers.py:93-125: ``boxes`` is a list of BoxList with equal lengths (:79); returns [bs, R, C, P, P] for several
        levels (:123) and, as the reference does (:104-106), the 4-D [bs*R, C, P, P] output of the only ROIAlign when
        there is a single level."""
        if len(x) != len(self.scales):
            raise ValueError(f"{len(self.scales)} poolers but {len(x)} feature levels")  # poolers.py:101-102
        counts = {len(b) for b in boxes}
        if len(counts) != 1:
            raise ValueError(f"all images must carry the same number of boxes, got {sorted(counts)}")  # :79
        rois = torch.stack([b.convert("xyxy").bbox for b in boxes], dim=0).to(torch.float32)
        out = self.forward_fixed(x, rois)
        if len(self.scales) == 1:
            return out.reshape(out.size(0) * out.size(1), out.size(2), out.size(3), out.size(4))
        return out


def make_pooler(cfg, head_name):
    """poolers.py:128-135."""
    resolution = cfg.MODEL[head_name].POOLER_RESOLUTION
    scales = cfg.MODEL[head_name].POOLER_SCALES
    sampling_ratio = cfg.MODEL[head_name].POOLER_SAMPLING_RATIO
    return Pooler(output_size=(resolution, resolution), scales=scales, sampling_ratio=sampling_ratio)
