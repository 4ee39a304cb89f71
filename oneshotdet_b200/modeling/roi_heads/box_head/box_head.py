"""Host side of the second stage's dense head (csrc/box_head.cu, tcgen05 tensor cores).

Mirrors what ``ROIBoxHead.forward`` does between its pooler and its post-processor
(maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:118-157, comparison_method 'concat', one support, no
negative support, LINEAR_FUSION off) plus ``FPNPredictor.forward`` (roi_box_predictors.py:80-84).  ``BoxHeadDense``
carries the reference's own sub-module names (``compress_dim_conv``, ``feature_aggreg``, ``fc6``, ``fc7``,
``predictor.cls_score``, ``predictor.bbox_pred``), so a ROIBoxHead state_dict loads unchanged; its parameters are
repacked once (bf16, K-major, the permutations the kernels' operand layouts need) and cached until one changes."""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from ...._lib import BoxHeadDesc, OsdError
from .... import _lib


class FPNPredictor(nn.Module):
    """roi_box_predictors.py:37-85 (parameters only; the forward lives in the GEMM kernel)."""

    def __init__(self, representation_size: int, num_classes: int = 2, num_bbox_reg_classes: int = 2):
        super().__init__()
        self.cls_score = nn.Linear(representation_size, num_classes)
        self.bbox_pred = nn.Linear(representation_size, num_bbox_reg_classes * 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)                 # :70-73
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in (self.cls_score, self.bbox_pred):
            nn.init.constant_(l.bias, 0)


class PackedBoxHeadWeights:
    """bf16 / permuted copies of the head's parameters in the layouts the kernels read."""

    def __init__(self, m: "BoxHeadDense", device):
        conv1, gn1, act1, conv2, gn2, act2 = list(m.compress_dim_conv)
        conv3, gn3, act3 = list(m.feature_aggreg)
        c2, c = conv1.out_channels, conv2.out_channels
        ch = conv3.out_channels
        if c2 != 2 * c or conv1.in_channels != c2 or conv2.in_channels != c2 or conv3.in_channels != c or 2 * ch != c or \
                conv1.kernel_size != (1, 1) or conv2.kernel_size != (1, 1) or conv3.kernel_size != (3, 3) or \
                conv3.padding != (1, 1):
            raise OsdError("box head: expected Conv1x1(2C,2C) GN LReLU Conv1x1(2C,C) GN LReLU and Conv3x3(C,C/2,pad 1) GN LReLU "
                           "(box_head.py:43-67)")
        if any(g.num_groups != 32 for g in (gn1, gn2, gn3)):
            raise OsdError("box head: GroupNorm must have 32 groups")
        if len({float(g.eps) for g in (gn1, gn2, gn3)}) != 1 or len({float(a.negative_slope) for a in (act1, act2, act3)}) != 1:
            raise OsdError("box head: the three GroupNorms / LeakyReLUs must share eps / slope")
        p = m.resolution
        mlp = m.fc6.out_features
        if m.fc6.in_features != ch * p * p or m.fc7.in_features != mlp or m.fc7.out_features != mlp:
            raise OsdError("box head: fc6 / fc7 shapes do not match (box_head.py:75-76)")
        self.channels, self.mlp, self.eps, self.slope = c, mlp, float(gn1.eps), float(act1.negative_slope)
        self.num_classes = m.predictor.cls_score.out_features
        self.num_box_out = m.predictor.bbox_pred.out_features
        f32 = dict(device=device, dtype=torch.float32)
        bf = lambda t: t.contiguous().to(torch.bfloat16)   # noqa: E731
        v = lambda t: t.detach().to(**f32).contiguous()    # noqa: E731
        self.w1 = bf(v(conv1.weight).reshape(c2, c2))
        self.w2 = bf(v(conv2.weight).reshape(c, c2))
        self.w3 = bf(v(conv3.weight).permute(0, 2, 3, 1).reshape(ch, 9 * c))                  # (out, ky, kx, in)
        self.w6 = bf(v(m.fc6.weight).reshape(mlp, ch, p * p).permute(0, 2, 1).reshape(mlp, p * p * ch))   # (out, pixel, channel)
        self.w7 = bf(v(m.fc7.weight))
        self.wp = bf(torch.cat((v(m.predictor.cls_score.weight), v(m.predictor.bbox_pred.weight)), 0))
        self.b1, self.b2, self.b3 = v(conv1.bias), v(conv2.bias), v(conv3.bias)
        self.g1w, self.g1b, self.g2w, self.g2b, self.g3w, self.g3b = (v(t) for t in (gn1.weight, gn1.bias, gn2.weight, gn2.bias,
                                                                                      gn3.weight, gn3.bias))
        self.b6, self.b7 = v(m.fc6.bias), v(m.fc7.bias)
        self.bp = torch.cat((v(m.predictor.cls_score.bias), v(m.predictor.bbox_pred.bias)), 0).contiguous()


class BoxHeadDense(nn.Module):
    """The dense middle of ROIBoxHead: ``forward(pooled, features_supp_roipooled)`` ->
    ``(class_logits [B*R, num_classes], box_regression [B*R, 8])``.

    pooled [B, R, C, 7, 7] (or [B*R, C, 7, 7] with ``batch_size``) is the Pooler's output, features_supp_roipooled
    [B, 1, C, 7, 7] the ROI-pooled support (box_head.py:118-131).  CUDA only: there is no CPU path."""

    def __init__(self, channels: int = 256, mlp_dim: int = 1024, resolution: int = 7, num_classes: int = 2,
                 roi_chunk: int = 0):
        super().__init__()
        c, c2 = channels, 2 * channels
        self.resolution, self.roi_chunk = resolution, roi_chunk
        self.compress_dim_conv = nn.Sequential(nn.Conv2d(c2, c2, 1), nn.GroupNorm(32, c2), nn.LeakyReLU(0.2),
                                               nn.Conv2d(c2, c, 1), nn.GroupNorm(32, c), nn.LeakyReLU(0.2))
        for l in self.compress_dim_conv:                                 # box_head.py:52-54
            if isinstance(l, nn.Conv2d):
                nn.init.normal_(l.weight, std=0.01)
        self.feature_aggreg = nn.Sequential(nn.Conv2d(c, c // 2, 3, 1, 1), nn.GroupNorm(32, c // 2), nn.LeakyReLU(0.2))
        self.fc6 = nn.Linear((c // 2) * resolution ** 2, mlp_dim)
        self.fc7 = nn.Linear(mlp_dim, mlp_dim)
        self.predictor = FPNPredictor(mlp_dim, num_classes, 2)
        self._packed = None
        self._ws = None     # grow-only scratch of the kernels (single-stream use, like PreparedFusion)

    def packed(self, device) -> PackedBoxHeadWeights:
        key = (str(device), tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, PackedBoxHeadWeights(self, device))
        return self._packed[1]

    @torch.no_grad()
    def forward(self, pooled, features_supp_roipooled, batch_size: int | None = None):
        """``pooled``: fp32 [B,R,C,7,7] / [B*R,C,7,7] (the reference layout), or bf16 [B,R,49,C] / [B*R,49,C] as
        ``Pooler.forward_fixed(..., rows_bf16=True)`` writes it (no repacking pass)."""
        rows = pooled.dtype == torch.bfloat16
        lead = 2 if pooled.dim() == (4 if rows else 5) else 1
        if lead == 2:
            b, r = pooled.shape[:2]
            pooled = pooled.reshape(b * r, *pooled.shape[2:])
        else:
            if batch_size is None:
                raise OsdError("box head: a pooled tensor without the [B, R] leading dimensions needs batch_size")
            b = int(batch_size)
            if pooled.size(0) % b != 0:
                raise OsdError("box head: rows of pooled must be a multiple of the batch size")
            r = pooled.size(0) // b
        dev = pooled.device
        _lib.require_device(dev)
        w = self.packed(dev)
        c, p = w.channels, self.resolution
        supp = features_supp_roipooled
        if supp.numel() != b * c * p * p:
            raise OsdError(f"box head: one support of [C={c},{p},{p}] per episode expected (box_head.py:120), got {tuple(supp.shape)}")
        want = (p * p, c) if rows else (c, p, p)
        if tuple(pooled.shape[1:]) != want or pooled.dtype not in (torch.float32, torch.bfloat16) or \
                supp.dtype != torch.float32 or supp.device != dev:
            raise OsdError(f"box head: pooled must be [B*R,{c},{p},{p}] float32 or [B*R,{p * p},{c}] bfloat16, the support float32 "
                           "on the same device")
        pooled, supp = pooled.contiguous(), supp.reshape(b, c, p, p).contiguous()
        n = b * r
        logits = torch.empty((n, w.num_classes), dtype=torch.float32, device=dev)
        reg = torch.empty((n, w.num_box_out), dtype=torch.float32, device=dev)
        d = BoxHeadDesc()
        d.batch, d.rois_per_image, d.channels, d.pooled_size = b, r, c, p
        d.mlp_dim, d.num_classes, d.num_box_out, d.roi_chunk = w.mlp, w.num_classes, w.num_box_out, int(self.roi_chunk)
        d.gn_eps, d.lrelu_slope = w.eps, w.slope
        if rows:
            d.pooled, d.pooled_nhwc_bf16 = None, pooled.data_ptr()
        else:
            d.pooled = pooled.data_ptr()
        d.supp = supp.data_ptr()
        d.w1, d.b1, d.gn1_w, d.gn1_b = w.w1.data_ptr(), w.b1.data_ptr(), w.g1w.data_ptr(), w.g1b.data_ptr()
        d.w2, d.b2, d.gn2_w, d.gn2_b = w.w2.data_ptr(), w.b2.data_ptr(), w.g2w.data_ptr(), w.g2b.data_ptr()
        d.w3, d.b3, d.gn3_w, d.gn3_b = w.w3.data_ptr(), w.b3.data_ptr(), w.g3w.data_ptr(), w.g3b.data_ptr()
        d.w6, d.b6, d.w7, d.b7 = w.w6.data_ptr(), w.b6.data_ptr(), w.w7.data_ptr(), w.b7.data_ptr()
        d.wp, d.bp = w.wp.data_ptr(), w.bp.data_ptr()
        d.class_logits, d.box_regression = logits.data_ptr(), reg.data_ptr()
        lib = _lib.load()
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.osd_box_head_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "osd_box_head_workspace_bytes")
        ws = self._ws
        if ws is None or ws.numel() < nbytes.value or ws.device != dev:
            ws = self._ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.osd_box_head_forward(ctypes.byref(d), ws.data_ptr(), ws.numel(), _lib.current_stream_ptr(dev))
        _lib.check(rc, "osd_box_head_forward")
        return logits, reg
