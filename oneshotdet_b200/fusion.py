"""Host side of the 1x1 fusion-conv matching mode (csrc/fusion_conv.cu, tcgen05 tensor cores).

``fusion_forward(features, supp_pooled, batch_size, compress_dim_conv)`` evaluates the reference's
``compress_dim_conv`` (modeling/roi_heads/box_head/box_head.py:43-54) on cat((x, support.expand_as(x)), 1)
(:147-149) for every FPN level with bf16 operands and fp32 accumulation.  The module is the stock
``nn.Sequential(Conv2d, GroupNorm, LeakyReLU, Conv2d, GroupNorm, LeakyReLU)``; its parameters are repacked once
(bf16 weight halves, transposed support half) and cached until a parameter changes."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import FusionDesc, OsdError, OSD_MAX_LEVELS

STAGES = {"conv1": 0, "full": 1}


class PackedFusionWeights:
    """bf16 / transposed copies of compress_dim_conv's parameters in the layout the kernels read."""

    def __init__(self, module, device):
        conv1, gn1, act1, conv2, gn2, act2 = list(module)
        c2 = conv1.out_channels
        c = c2 // 2
        if conv1.in_channels != c2 or conv2.in_channels != c2 or conv2.out_channels != c or \
                conv1.kernel_size != (1, 1) or conv2.kernel_size != (1, 1):
            raise OsdError("fusion: expected Conv2d(2C,2C,1) ... Conv2d(2C,C,1) (box_head.py:43-51)")
        if gn1.num_groups != 32 or gn2.num_groups != 32:
            raise OsdError("fusion: GroupNorm must have 32 groups (box_head.py:46,49)")
        self.channels = c
        self.eps = float(gn1.eps)
        if float(gn2.eps) != self.eps:
            raise OsdError("fusion: both GroupNorms must share eps")
        self.slope = float(act1.negative_slope)
        f32 = dict(device=device, dtype=torch.float32)
        w1 = conv1.weight.detach().to(**f32).reshape(c2, c2)
        self.w1x = w1[:, :c].contiguous().to(torch.bfloat16)          # [2C, C]  target half, K-major
        self.w1s_t = w1[:, c:].t().contiguous()                       # [C, 2C]  support half, folded into a bias
        self.b1 = (conv1.bias.detach() if conv1.bias is not None else torch.zeros(c2)).to(**f32).contiguous()
        self.gn1_w = gn1.weight.detach().to(**f32).contiguous()
        self.gn1_b = gn1.bias.detach().to(**f32).contiguous()
        self.w2 = conv2.weight.detach().to(**f32).reshape(c, c2).contiguous().to(torch.bfloat16)   # [C, 2C]
        self.b2 = (conv2.bias.detach() if conv2.bias is not None else torch.zeros(c)).to(**f32).contiguous()
        self.gn2_w = gn2.weight.detach().to(**f32).contiguous()
        self.gn2_b = gn2.bias.detach().to(**f32).contiguous()
        # M_g = sum over the channels of GroupNorm-1 group g of w_c w_c^T (w = the bf16-rounded target half): lets pass A
        # take the GN-1 statistics from a Gram GEMM of the activations instead of running conv1 (csrc/fusion_fused.cu)
        self.w1x_gram = None
        if c >= 128:
            wg = self.w1x.float().reshape(32, c2 // 32, c)
            self.w1x_gram = torch.einsum("gck,gcl->gkl", wg, wg).contiguous()
        self.versions = tuple(p._version for p in module.parameters())


def packed_weights(module, device) -> PackedFusionWeights:
    cache = getattr(module, "_osd_packed", None)
    key = (str(device), tuple(p._version for p in module.parameters()), tuple(p.data_ptr() for p in module.parameters()))
    if cache is None or cache[0] != key:
        cache = (key, PackedFusionWeights(module, device))
        module._osd_packed = cache
    return cache[1]


class PreparedFusion:
    """A fixed-shape fusion call marshalled once (descriptor, packed weights, outputs, workspace); ``__call__`` is one
    C-ABI invocation: bias kernel + two persistent tcgen05 GEMM launches + the final GroupNorm/LeakyReLU pass."""

    def __init__(self, features, supp_pooled, batch_size: int, module, stage: str = "full", out=None):
        if stage not in STAGES:
            raise OsdError(f"fusion_forward: unknown stage '{stage}'")
        self.lib = _lib.load()
        nl = len(features)
        if nl == 0 or nl > OSD_MAX_LEVELS or len(supp_pooled) != nl:
            raise OsdError("fusion_forward: need 1..8 levels and one support tensor per level")
        dev = features[0].device
        _lib.require_device(dev)
        b, c = features[0].shape[:2]
        if b != batch_size:
            raise OsdError(f"fusion_forward: batch_size {batch_size} does not match features batch {b}")
        w = packed_weights(module, dev)
        if w.channels != c:
            raise OsdError(f"fusion_forward: module is built for C={w.channels}, features have C={c}")
        d = FusionDesc()
        d.num_levels, d.batch, d.channels, d.stage = nl, b, c, STAGES[stage]
        d.gn_eps, d.lrelu_slope = w.eps, w.slope
        cout = 2 * c if stage == "conv1" else c
        self.outs, self.inputs = [], []
        shots = None
        for l, (f, s) in enumerate(zip(features, supp_pooled)):
            if f.dim() != 4 or f.size(0) != b or f.size(1) != c or f.dtype != torch.float32 or f.device != dev:
                raise OsdError(f"fusion_forward: level {l}: features must be [B={b},C={c},H,W] float32 on {dev}")
            if s.dtype != torch.float32 or s.device != dev or s.numel() == 0 or s.numel() % (b * c) != 0:
                raise OsdError(f"fusion_forward: level {l}: support must be [B*S,{c},1,1] float32")
            sl = s.numel() // (b * c)
            shots = sl if shots is None else shots
            if sl != shots:
                raise OsdError("fusion_forward: every level must carry the same number of shots")
            f = f.contiguous()
            s = s.reshape(b * sl, c).contiguous()
            h, wd = f.shape[-2:]
            o = out[l] if out is not None else torch.empty((b, cout, h, wd), dtype=torch.float32, device=dev)
            if tuple(o.shape) != (b, cout, h, wd) or o.dtype != torch.float32 or not o.is_contiguous():
                raise OsdError(f"fusion_forward: out[{l}] has the wrong shape, dtype or layout")
            d.hw[l] = h * wd
            d.feat[l], d.supp[l], d.out[l] = f.data_ptr(), s.data_ptr(), o.data_ptr()
            self.inputs += [f, s]
            self.outs.append(o)
        d.shots = shots
        d.w1x_bf16, d.w1s_t, d.b1 = w.w1x.data_ptr(), w.w1s_t.data_ptr(), w.b1.data_ptr()
        d.gn1_w, d.gn1_b = w.gn1_w.data_ptr(), w.gn1_b.data_ptr()
        d.w2_bf16, d.b2 = w.w2.data_ptr(), w.b2.data_ptr()
        d.gn2_w, d.gn2_b = w.gn2_w.data_ptr(), w.gn2_b.data_ptr()
        d.w1x_gram = w.w1x_gram.data_ptr() if w.w1x_gram is not None else None
        nbytes = ctypes.c_size_t(0)
        _lib.check(self.lib.osd_fusion_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)),
                   "osd_fusion_workspace_bytes")
        self.ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        self.desc, self.weights, self.device = d, w, dev
        self.locations = sum(int(d.hw[l]) for l in range(nl))
        self.batch, self.channels = b, c

    def __call__(self):
        with torch.cuda.device(self.device):
            rc = self.lib.osd_fusion_forward(ctypes.byref(self.desc), self.ws.data_ptr(), self.ws.numel(),
                                             _lib.current_stream_ptr(self.device))
        _lib.check(rc, "osd_fusion_forward")
        return self.outs


@torch.no_grad()
def fusion_forward(features, supp_pooled, batch_size: int, module, stage: str = "full", out=None):
    """features[l] [B,C,H,W] fp32 NCHW, supp_pooled[l] [B*S,C,1,1] fp32 -> list of [B,C,H,W] (stage 'full') or the
    first convolution's [B,2C,H,W] (stage 'conv1').  All levels and episodes go through one persistent GEMM
    launch per convolution."""
    return PreparedFusion(list(features), list(supp_pooled), batch_size, module, stage, out)()
