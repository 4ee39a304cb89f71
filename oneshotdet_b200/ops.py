"""Functional host API over libosd_b200.so: tensors in, tensors out, everything enqueued on the current
CUDA stream.  These are the calls the reference-shaped wrappers (layers.nms, FCOSPostProcessor,
MatchingModule) make; none of them has a CPU or eager-PyTorch fallback."""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import FcosConfig, FcosPlan, MatchDesc, NmsPlan, OsdError, OSD_MAX_LEVELS

MATCH_MODES = {"product": 0, "concat": 1, "concat_reversed": 2}

_workspace = _lib.Workspace()
_seg_cache: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
_hw_cache: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()


def _cached_small_tensor(cache, key, values, dtype, device):
    t = cache.get(key)
    if t is None:
        t = torch.tensor(values, dtype=dtype, device=device)
        cache[key] = t
        if len(cache) > 256:
            cache.popitem(last=False)
    else:
        cache.move_to_end(key)
    return t


def launch_count() -> int:
    return int(_lib.load().osd_launch_count())


def reset_launch_count() -> None:
    _lib.load().osd_reset_launch_count()


# --------------------------------------------------------------------------------------------------
# NMS  (reference: maskrcnn_benchmark/csrc/nms.h:10-28, csrc/cpu/nms_cpu.cpp:5-75, csrc/cuda/nms.cu:70-131)
# --------------------------------------------------------------------------------------------------
def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, seg_offsets: torch.Tensor, max_seg_len: int,
                threshold: float, strict: bool = False):
    """E independent NMS problems in one launch sequence.

    boxes [N,4] fp32 xyxy, scores [N] fp32, seg_offsets int64 [E+1] on the same CUDA device;
    ``max_seg_len`` is a host-side upper bound of every segment length.
    Returns (keep int64 [N], counts int32 [E]); segment e's kept rows are the ascending global row
    indices ``keep[seg[e] : seg[e] + counts[e]]``.  No host synchronisation."""
    _lib.require_device(boxes.device)
    if boxes.dtype != torch.float32 or scores.dtype != torch.float32:
        raise OsdError("batched_nms: boxes and scores must be float32 (as nms_cuda, csrc/cuda/nms.cu:71)")
    if boxes.dim() != 2 or boxes.size(1) != 4 or scores.dim() != 1 or scores.size(0) != boxes.size(0):
        raise OsdError(f"batched_nms: expected boxes [N,4] and scores [N], got {tuple(boxes.shape)} / {tuple(scores.shape)}")
    if seg_offsets.dtype != torch.int64 or seg_offsets.device != boxes.device or seg_offsets.dim() != 1:
        raise OsdError("batched_nms: seg_offsets must be an int64 vector on the boxes' device")
    lib = _lib.load()
    boxes = boxes.contiguous()
    scores = scores.contiguous()
    seg_offsets = seg_offsets.contiguous()
    num_seg = seg_offsets.numel() - 1
    n = boxes.size(0)
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    counts = torch.zeros(max(num_seg, 0), dtype=torch.int32, device=boxes.device)
    if num_seg <= 0:
        return keep, counts
    plan = NmsPlan()
    _lib.check(lib.osd_batched_nms_plan(num_seg, int(max_seg_len), ctypes.byref(plan)), "osd_batched_nms_plan")
    ws = _workspace.get(boxes.device, plan.workspace_bytes)
    with torch.cuda.device(boxes.device):
        rc = lib.osd_batched_nms(boxes.data_ptr(), scores.data_ptr(), seg_offsets.data_ptr(), num_seg, int(max_seg_len),
                                 float(threshold), int(bool(strict)), ws.data_ptr(), ws.numel(), keep.data_ptr(),
                                 counts.data_ptr(), _lib.current_stream_ptr(boxes.device))
    _lib.check(rc, "osd_batched_nms")
    return keep, counts


def nms(dets: torch.Tensor, scores: torch.Tensor, threshold: float, strict: bool = False) -> torch.Tensor:
    """Drop-in for ``maskrcnn_benchmark._C.nms`` (csrc/vision.cpp:8, csrc/nms.h:10-28) on CUDA tensors.

    Returns the kept ORIGINAL indices in ascending order, int64, on the input device (csrc/cuda/nms.cu:127-130).
    Suppression uses ``IoU >= threshold`` like the runnable reference path nms_cpu (csrc/cpu/nms_cpu.cpp:60);
    ``strict=True`` selects the reference CUDA kernel's ``>`` (csrc/cuda/nms.cu:60).
    Empty input returns an empty int64 *CPU* tensor, as the reference dispatcher does (csrc/nms.h:17-18)."""
    if not dets.is_cuda:
        raise OsdError("oneshotdet_b200.nms: dets must be a CUDA tensor; this package has no CPU path "
                       "(the reference's nms_cpu lives in the oracle, for tests only)")
    if not scores.is_cuda:
        raise OsdError("oneshotdet_b200.nms: scores must be a CUDA tensor")
    if dets.dtype != scores.dtype:
        raise OsdError("dets should have the same type as scores")  # nms_cpu.cpp:11
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device="cpu")
    if dets.dtype != torch.float32:
        raise OsdError("oneshotdet_b200.nms: only float32 is supported on CUDA (csrc/cuda/nms.cu:71)")
    n = dets.size(0)
    seg = _cached_small_tensor(_seg_cache, (dets.device.index, n), [0, n], torch.int64, dets.device)
    keep, counts = batched_nms(dets, scores.reshape(-1), seg, n, threshold, strict)
    k = int(counts[0].item())  # dynamic output shape: the one host sync of this call
    return keep[:k]


# --------------------------------------------------------------------------------------------------
# FCOS post-processing  (reference: modeling/rpn/fcos/inference.py:46-137, :251-323)
# --------------------------------------------------------------------------------------------------
@dataclass
class FcosResult:
    boxes: torch.Tensor        # [B, K, 4] fp32
    scores: torch.Tensor       # [B, K]
    index: torch.Tensor        # [B, K] int32 compact candidate index of each row
    count: torch.Tensor        # [B] int32 valid rows per episode
    plan: FcosPlan
    workspace: torch.Tensor    # keeps the intermediates alive for inspection
    block: torch.Tensor = None  # uint8: the ONE allocation boxes | scores | index | count are views of (multi-GPU payload)

    def _view(self, off, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off:off + nbytes].view(dtype).view(*shape)

    def candidates(self):
        """(cand_boxes [B,CAP,4], cand_scores [B,CAP], cand_loc [B,CAP], level_count [B,L], level_slot list):
        the slotted candidate arrays the NMS stage consumed."""
        b = self.boxes.size(0)
        cap = self.plan.cand_capacity
        nl = self.num_levels
        return (self._view(self.plan.off_cand_boxes, torch.float32, (b, cap, 4)),
                self._view(self.plan.off_cand_scores, torch.float32, (b, cap)),
                self._view(self.plan.off_cand_loc, torch.int32, (b, cap)),
                self._view(self.plan.off_level_count, torch.int32, (b, nl)),
                [self.plan.level_slot[l] for l in range(nl)])

    def kept_before_cut(self):
        return self._view(self.plan.off_kept_count, torch.int32, (self.boxes.size(0),))

    num_levels: int = 0


def result_block_layout(b: int, k: int):
    """Byte offsets of boxes fp32 [b,k,4] | scores fp32 [b,k] | index int32 [b,k] | count int32 [b] inside a result block
    (each part 256-byte aligned) and the block size."""
    def up(x):
        return (x + 255) // 256 * 256
    o_boxes = 0
    o_scores = up(o_boxes + b * k * 16)
    o_index = up(o_scores + b * k * 4)
    o_count = up(o_index + b * k * 4)
    return (o_boxes, o_scores, o_index, o_count), up(o_count + b * 4)


def result_block(b: int, k: int, device, block: torch.Tensor | None = None):
    """(block uint8, (boxes, scores, index, count) views).  ``block`` given: views of an existing (e.g. gathered) block."""
    (o_boxes, o_scores, o_index, o_count), size = result_block_layout(b, k)
    if block is None:
        block = torch.zeros(size, dtype=torch.uint8, device=device)
    def view(off, nbytes, dtype, shape):
        return block[off:off + nbytes].view(dtype).view(*shape)
    return block, (view(o_boxes, b * k * 16, torch.float32, (b, k, 4)), view(o_scores, b * k * 4, torch.float32, (b, k)),
                   view(o_index, b * k * 4, torch.int32, (b, k)), view(o_count, b * 4, torch.int32, (b,)))


def fcos_config(level_shapes, strides, batch, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_nms_top_n, min_size,
                strict=False, early_exit=True, reg_scales=None) -> FcosConfig:
    if len(level_shapes) > OSD_MAX_LEVELS:
        raise OsdError(f"at most {OSD_MAX_LEVELS} FPN levels are supported")
    cfg = FcosConfig()
    cfg.num_levels = len(level_shapes)
    cfg.batch = int(batch)
    for l, ((h, w), s) in enumerate(zip(level_shapes, strides)):
        cfg.height[l], cfg.width[l], cfg.stride[l] = int(h), int(w), int(s)
    cfg.pre_nms_thresh = float(pre_nms_thresh)
    cfg.pre_nms_top_n = int(pre_nms_top_n)
    cfg.nms_thresh = float(nms_thresh)
    cfg.post_nms_top_n = int(post_nms_top_n)
    cfg.min_size = float(min_size)
    cfg.strict = int(bool(strict))
    cfg.early_exit = int(bool(early_exit))
    if reg_scales is not None:  # reg[l] is the raw bbox_pred conv output: exp(x * scale_l) is evaluated in the kernel
        if len(reg_scales) != len(level_shapes):
            raise OsdError(f"fcos_postprocess: {len(level_shapes)} levels but {len(reg_scales)} regression scales")
        cfg.reg_transform = 1
        for l, v in enumerate(reg_scales):
            cfg.reg_scale[l] = float(v)
    return cfg


class PreparedFcos:
    """A fixed-shape post-processing call with every host-side argument marshalled once: ``__call__`` is a single
    C-ABI invocation on the current stream (serving loops, CUDA-graph capture)."""

    def __init__(self, cls, reg, ctr, strides, image_sizes, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_nms_top_n,
                 min_size=0.0, strict=False, early_exit=True, workspace=None, private_workspace=False, reg_scales=None):
        self.lib = _lib.load()
        dev = cls[0].device
        _lib.require_device(dev)
        b = cls[0].size(0)
        shapes = []
        for l, (c, r, t) in enumerate(zip(cls, reg, ctr)):
            if c.dtype != torch.float32 or r.dtype != torch.float32 or t.dtype != torch.float32:
                raise OsdError("fcos_postprocess: head outputs must be float32")
            h, w = c.shape[-2:]
            if tuple(c.shape) != (b, 1, h, w) or tuple(r.shape) != (b, 4, h, w) or tuple(t.shape) != (b, 1, h, w):
                raise OsdError(f"fcos_postprocess: level {l}: expected cls [B,1,H,W], reg [B,4,H,W], ctr [B,1,H,W]; got "
                               f"{tuple(c.shape)}, {tuple(r.shape)}, {tuple(t.shape)}")
            shapes.append((h, w))
        self.cls = [c.contiguous() for c in cls]
        self.reg = [r.contiguous() for r in reg]
        self.ctr = [t.contiguous() for t in ctr]
        self.device = dev
        self.cfg = fcos_config(shapes, strides, b, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_nms_top_n, min_size,
                               strict, early_exit, reg_scales)
        self.plan = FcosPlan()
        _lib.check(self.lib.osd_fcos_postprocess_plan(ctypes.byref(self.cfg), ctypes.byref(self.plan)),
                   "osd_fcos_postprocess_plan")
        if isinstance(image_sizes, torch.Tensor):
            hw = image_sizes.to(device=dev, dtype=torch.int32).contiguous()
        else:
            key = (dev.index, tuple((int(h), int(w)) for h, w in image_sizes))
            hw = _cached_small_tensor(_hw_cache, key, [[int(h), int(w)] for h, w in image_sizes], torch.int32, dev)
        if hw.numel() != 2 * b:
            raise OsdError(f"fcos_postprocess: {b} episodes but {hw.numel() // 2} image sizes")
        self.hw = hw
        k = self.plan.out_capacity
        # one allocation for the four outputs, so that a multi-GPU gather can ship the step's result as a single buffer
        block, (out_boxes, out_scores, out_index, out_count) = result_block(b, k, dev)
        if workspace is not None:
            ws = workspace
        elif private_workspace:  # results (candidates, kept counts) survive unrelated calls
            ws = torch.empty(self.plan.workspace_bytes, dtype=torch.uint8, device=dev)
        else:
            ws = _workspace.get(dev, self.plan.workspace_bytes)
        self.result = FcosResult(out_boxes, out_scores, out_index, out_count, self.plan, ws, num_levels=len(shapes),
                                 block=block)
        nl = len(shapes)
        arr = ctypes.c_void_p * nl
        self._cls = arr(*[c.data_ptr() for c in self.cls])
        self._reg = arr(*[r.data_ptr() for r in self.reg])
        self._ctr = arr(*[t.data_ptr() for t in self.ctr])
        self.batch = b
        self.launches = None

    def __call__(self) -> FcosResult:
        if self.batch == 0:
            return self.result
        r = self.result
        with torch.cuda.device(self.device):
            rc = self.lib.osd_fcos_postprocess(ctypes.byref(self.cfg), self._cls, self._reg, self._ctr, self.hw.data_ptr(),
                                               r.workspace.data_ptr(), r.workspace.numel(), r.boxes.data_ptr(),
                                               r.scores.data_ptr(), r.index.data_ptr(), r.count.data_ptr(),
                                               _lib.current_stream_ptr(self.device))
        _lib.check(rc, "osd_fcos_postprocess")
        return r


def fcos_postprocess(cls, reg, ctr, strides, image_sizes, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_nms_top_n,
                     min_size=0.0, strict=False, early_exit=True, workspace: torch.Tensor | None = None,
                     reg_scales=None) -> FcosResult:
    """Fused score / top-k / decode / clip / NMS / post-top-n for all levels and episodes.

    cls[l] [B,1,H,W] logits, reg[l] [B,4,H,W] ltrb distances, ctr[l] [B,1,H,W] logits (fp32 CUDA, NCHW);
    image_sizes: list of (h, w) per episode, or an int32 CUDA tensor [B,2].  No host synchronisation.
    reg_scales (one float per level): reg[l] is the RAW output of the head's bbox_pred conv and the head's tail
    ``torch.exp(scales[l](x))`` (fcos.py:95-97) is folded into the decode of the selected locations."""
    return PreparedFcos(cls, reg, ctr, strides, image_sizes, pre_nms_thresh, pre_nms_top_n, nms_thresh, post_nms_top_n,
                        min_size, strict, early_exit, workspace, reg_scales=reg_scales)()


# --------------------------------------------------------------------------------------------------
# matching  (reference: modeling/detector/generalized_rcnn.py:100-104, :306-311; box_head.py:144-147)
# --------------------------------------------------------------------------------------------------
class PreparedMatch:
    """A fixed-shape matching call marshalled once; ``__call__`` is one C-ABI invocation (one kernel launch)."""

    def __init__(self, features, supp_pooled, batch_size: int, mode: str = "product", out=None):
        if mode not in MATCH_MODES:
            raise OsdError(f"match_forward: unknown mode '{mode}' (expected one of {sorted(MATCH_MODES)})")
        self.lib = _lib.load()
        nl = len(features)
        if nl == 0 or nl > OSD_MAX_LEVELS or len(supp_pooled) != nl:
            raise OsdError("match_forward: need 1..8 levels and one support tensor per level")
        dev = features[0].device
        _lib.require_device(dev)
        dtype = features[0].dtype
        if dtype not in (torch.float32, torch.bfloat16):
            raise OsdError("match_forward: features must be float32 or bfloat16")
        b, c = features[0].shape[:2]
        if b != batch_size:
            raise OsdError(f"match_forward: batch_size {batch_size} does not match features batch {b}")
        channels_last = features[0].dim() == 4 and not features[0].is_contiguous() and \
            features[0].is_contiguous(memory_format=torch.channels_last)
        d = MatchDesc()
        d.num_levels, d.batch, d.channels = nl, b, c
        d.mode = MATCH_MODES[mode]
        d.layout = 1 if channels_last else 0
        d.dtype = 0 if dtype == torch.float32 else 1
        cout = c if mode == "product" else 2 * c
        self.outs, self.inputs = [], []
        shots = None
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        for l, (f, s) in enumerate(zip(features, supp_pooled)):
            if f.dim() != 4 or f.size(0) != b or f.size(1) != c or f.dtype != dtype or f.device != dev:
                raise OsdError(f"match_forward: level {l}: features must all be [B={b},C={c},H,W] {dtype} on {dev}")
            if s.dtype != dtype or s.device != dev or s.numel() % max(b * c, 1) != 0 or (s.numel() == 0 and b > 0):
                raise OsdError(f"match_forward: level {l}: support must be [B*S,{c},1,1] {dtype}, got {tuple(s.shape)}")
            sl = s.numel() // (b * c) if b > 0 else 1
            shots = sl if shots is None else shots
            if sl != shots:
                raise OsdError("match_forward: every level must carry the same number of shots")
            f = f.contiguous(memory_format=fmt)
            s = s.reshape(b * sl, c).contiguous()
            h, w = f.shape[-2:]
            if out is not None:
                o = out[l]
                if tuple(o.shape) != (b, cout, h, w) or o.dtype != dtype or not o.is_contiguous(memory_format=fmt):
                    raise OsdError(f"match_forward: out[{l}] has the wrong shape, dtype or memory format")
            else:
                o = torch.empty((b, cout, h, w), dtype=dtype, device=dev, memory_format=fmt)
            d.hw[l] = h * w
            d.feat[l], d.supp[l], d.out[l] = f.data_ptr(), s.data_ptr(), o.data_ptr()
            self.inputs += [f, s]
            self.outs.append(o)
        d.shots = shots
        self.desc = d
        self.device = dev
        self.batch = b

    def __call__(self):
        if self.batch == 0:
            return self.outs
        with torch.cuda.device(self.device):
            rc = self.lib.osd_match_forward(ctypes.byref(self.desc), _lib.current_stream_ptr(self.device))
        _lib.check(rc, "osd_match_forward")
        return self.outs


def match_forward(features, supp_pooled, batch_size: int, mode: str = "product", out=None):
    """features[l] [B,C,H,W] (NCHW-contiguous or channels_last; fp32 or bf16), supp_pooled[l] [B*S,C,1,1]
    (episode-major, shot-minor).  Returns a list of [B,C,H,W] (product) or [B,2C,H,W] (concat) tensors in the
    input's memory format.  All levels go out in one kernel launch."""
    return PreparedMatch(list(features), list(supp_pooled), batch_size, mode, out)()


# --------------------------------------------------------------------------------------------------
# second-stage box post-processing  (reference: modeling/roi_heads/box_head/inference.py:46-167,
# modeling/box_coder.py:52-95) -- SURVEY section 8(f) row 2
# --------------------------------------------------------------------------------------------------
SCORE_MODES = {"softmax": 0, "sigmoid": 1}
BBOX_XFORM_CLIP = 4.135166556742356  # math.log(1000. / 16), box_coder.py:21


@dataclass
class BoxPostResult:
    boxes: torch.Tensor        # [B, K, 4]
    scores: torch.Tensor       # [B, K]
    index: torch.Tensor        # [B, K] int32 compact candidate index
    count: torch.Tensor        # [B] int32
    plan: "_lib.BoxPostPlan"
    workspace: torch.Tensor

    def _view(self, off, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off:off + nbytes].view(dtype).view(*shape)

    def candidates(self):
        """(cand_boxes [B,R,4], cand_scores [B,R], cand_src [B,R] proposal rows, cand_count [B])."""
        b, r = self.boxes.size(0), self.plan.cand_capacity
        return (self._view(self.plan.off_cand_boxes, torch.float32, (b, r, 4)),
                self._view(self.plan.off_cand_scores, torch.float32, (b, r)),
                self._view(self.plan.off_cand_src, torch.int32, (b, r)),
                self._view(self.plan.off_cand_count, torch.int32, (b,)))


def box_postprocess(class_logits, box_regression, proposals, image_sizes, score_thresh=0.0, nms_thresh=0.5,
                    detections_per_img=2000, weights=(10.0, 10.0, 5.0, 5.0), score_mode="softmax", reg_offset=4,
                    roi_count=None, bbox_xform_clip=BBOX_XFORM_CLIP, strict=False, early_exit=True,
                    private_workspace=False) -> BoxPostResult:
    """Second-stage post-processing for all images of the batch, no host synchronisation.

    class_logits [B*R, L], box_regression [B*R, >= reg_offset+4] (fp32 CUDA), proposals [B, R, 4] xyxy,
    image_sizes list of (h, w) or int32 CUDA [B,2]; roi_count optional int32 CUDA [B] (valid rows per image)."""
    dev = class_logits.device
    _lib.require_device(dev)
    lib = _lib.load()
    if score_mode not in SCORE_MODES:
        raise OsdError(f"box_postprocess: unknown score_mode '{score_mode}' (expected one of {sorted(SCORE_MODES)})")
    if proposals.dim() != 3 or proposals.size(2) != 4:
        raise OsdError(f"box_postprocess: proposals must be [B,R,4], got {tuple(proposals.shape)}")
    b, r = proposals.size(0), proposals.size(1)
    for name, t in (("class_logits", class_logits), ("box_regression", box_regression), ("proposals", proposals)):
        if t.dtype != torch.float32 or t.device != dev:
            raise OsdError(f"box_postprocess: {name} must be float32 on {dev}")
    if class_logits.dim() != 2 or class_logits.size(0) != b * r or box_regression.dim() != 2 or box_regression.size(0) != b * r:
        raise OsdError(f"box_postprocess: expected {b * r} rows of logits / regression, got "
                       f"{tuple(class_logits.shape)} / {tuple(box_regression.shape)}")
    cfg = _lib.BoxPostConfig()
    cfg.batch, cfg.rois_per_image = b, max(r, 1)
    cfg.num_logits, cfg.reg_columns, cfg.reg_offset = class_logits.size(1), box_regression.size(1), int(reg_offset)
    cfg.score_mode = SCORE_MODES[score_mode]
    for k in range(4):
        cfg.weights[k] = float(weights[k])
    cfg.bbox_xform_clip = float(bbox_xform_clip)
    cfg.score_thresh, cfg.nms_thresh = float(score_thresh), float(nms_thresh)
    cfg.detections_per_img = int(detections_per_img)
    cfg.strict, cfg.early_exit = int(bool(strict)), int(bool(early_exit))
    plan = _lib.BoxPostPlan()
    _lib.check(lib.osd_box_postprocess_plan(ctypes.byref(cfg), ctypes.byref(plan)), "osd_box_postprocess_plan")
    if isinstance(image_sizes, torch.Tensor):
        hw = image_sizes.to(device=dev, dtype=torch.int32).contiguous()
    else:
        key = (dev.index, tuple((int(h), int(w)) for h, w in image_sizes))
        hw = _cached_small_tensor(_hw_cache, key, [[int(h), int(w)] for h, w in image_sizes], torch.int32, dev)
    if hw.numel() != 2 * b:
        raise OsdError(f"box_postprocess: {b} images but {hw.numel() // 2} image sizes")
    if roi_count is not None:
        if roi_count.dtype != torch.int32 or roi_count.device != dev or roi_count.numel() != b:
            raise OsdError("box_postprocess: roi_count must be an int32 vector [B] on the inputs' device")
        roi_count = roi_count.contiguous()
    k = plan.out_capacity
    out_boxes = torch.empty((b, k, 4), dtype=torch.float32, device=dev)
    out_scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    out_index = torch.empty((b, k), dtype=torch.int32, device=dev)
    out_count = torch.zeros((b,), dtype=torch.int32, device=dev)
    ws = (torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev) if private_workspace
          else _workspace.get(dev, plan.workspace_bytes))
    res = BoxPostResult(out_boxes, out_scores, out_index, out_count, plan, ws)
    if b == 0 or r == 0:
        return res
    cl, br, pr = class_logits.contiguous(), box_regression.contiguous(), proposals.contiguous()
    with torch.cuda.device(dev):
        rc = lib.osd_box_postprocess(ctypes.byref(cfg), cl.data_ptr(), br.data_ptr(), pr.data_ptr(),
                                     roi_count.data_ptr() if roi_count is not None else None, hw.data_ptr(),
                                     ws.data_ptr(), ws.numel(), out_boxes.data_ptr(), out_scores.data_ptr(),
                                     out_index.data_ptr(), out_count.data_ptr(), _lib.current_stream_ptr(dev))
    _lib.check(rc, "osd_box_postprocess")
    return res
