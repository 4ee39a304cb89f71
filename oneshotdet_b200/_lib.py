"""ctypes binding of libosd_b200.so (include/osd_b200.h).  No fallback: if the library is missing or
the device is not sm_100, every operator raises."""
from __future__ import annotations

import ctypes
import os
import threading

OSD_MAX_LEVELS = 8
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSD_B200_LIB") or os.path.join(_HERE, "lib", "libosd_b200.so")  # env: an alternative build

c_f32p = ctypes.c_void_p  # device pointers travel as integers
c_void_p = ctypes.c_void_p


class NmsPlan(ctypes.Structure):
    _fields_ = [("workspace_bytes", ctypes.c_size_t), ("padded_len", ctypes.c_int32), ("mask_words", ctypes.c_int32)]


class FcosConfig(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("batch", ctypes.c_int32),
                ("height", ctypes.c_int32 * OSD_MAX_LEVELS), ("width", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("stride", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("pre_nms_thresh", ctypes.c_float), ("pre_nms_top_n", ctypes.c_int32),
                ("nms_thresh", ctypes.c_float), ("post_nms_top_n", ctypes.c_int32),
                ("min_size", ctypes.c_float), ("strict", ctypes.c_int32), ("early_exit", ctypes.c_int32),
                ("reg_transform", ctypes.c_int32), ("reg_scale", ctypes.c_float * OSD_MAX_LEVELS)]


class FcosPlan(ctypes.Structure):
    _fields_ = [("workspace_bytes", ctypes.c_size_t), ("cand_capacity", ctypes.c_int32),
                ("out_capacity", ctypes.c_int32), ("level_slot", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("off_cand_boxes", ctypes.c_size_t), ("off_cand_scores", ctypes.c_size_t),
                ("off_cand_loc", ctypes.c_size_t), ("off_level_count", ctypes.c_size_t),
                ("off_kept_count", ctypes.c_size_t)]


class BoxPostConfig(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("rois_per_image", ctypes.c_int32), ("num_logits", ctypes.c_int32),
                ("reg_columns", ctypes.c_int32), ("reg_offset", ctypes.c_int32), ("score_mode", ctypes.c_int32),
                ("weights", ctypes.c_float * 4), ("bbox_xform_clip", ctypes.c_float), ("score_thresh", ctypes.c_float),
                ("nms_thresh", ctypes.c_float), ("detections_per_img", ctypes.c_int32), ("strict", ctypes.c_int32),
                ("early_exit", ctypes.c_int32)]


class BoxPostPlan(ctypes.Structure):
    _fields_ = [("workspace_bytes", ctypes.c_size_t), ("cand_capacity", ctypes.c_int32),
                ("out_capacity", ctypes.c_int32), ("off_cand_boxes", ctypes.c_size_t),
                ("off_cand_scores", ctypes.c_size_t), ("off_cand_src", ctypes.c_size_t),
                ("off_cand_count", ctypes.c_size_t), ("off_kept_count", ctypes.c_size_t)]


class MatchDesc(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("batch", ctypes.c_int32), ("shots", ctypes.c_int32),
                ("channels", ctypes.c_int32), ("mode", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("dtype", ctypes.c_int32), ("hw", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("feat", ctypes.c_void_p * OSD_MAX_LEVELS), ("supp", ctypes.c_void_p * OSD_MAX_LEVELS),
                ("out", ctypes.c_void_p * OSD_MAX_LEVELS)]


class FusionDesc(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("batch", ctypes.c_int32), ("shots", ctypes.c_int32),
                ("channels", ctypes.c_int32), ("stage", ctypes.c_int32), ("gn_eps", ctypes.c_float),
                ("lrelu_slope", ctypes.c_float), ("hw", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("feat", ctypes.c_void_p * OSD_MAX_LEVELS), ("supp", ctypes.c_void_p * OSD_MAX_LEVELS),
                ("out", ctypes.c_void_p * OSD_MAX_LEVELS),
                ("w1x_bf16", ctypes.c_void_p), ("w1s_t", ctypes.c_void_p), ("b1", ctypes.c_void_p),
                ("gn1_w", ctypes.c_void_p), ("gn1_b", ctypes.c_void_p), ("w2_bf16", ctypes.c_void_p),
                ("b2", ctypes.c_void_p), ("gn2_w", ctypes.c_void_p), ("gn2_b", ctypes.c_void_p),
                ("w1x_gram", ctypes.c_void_p)]


class SupportPoolDesc(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("num_supports", ctypes.c_int32), ("channels", ctypes.c_int32),
                ("mode", ctypes.c_int32), ("sampling_ratio", ctypes.c_int32),
                ("height", ctypes.c_int32 * OSD_MAX_LEVELS), ("width", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("spatial_scale", ctypes.c_float * OSD_MAX_LEVELS),
                ("feat", ctypes.c_void_p * OSD_MAX_LEVELS), ("out", ctypes.c_void_p * OSD_MAX_LEVELS),
                ("rois", ctypes.c_void_p)]


class RoiPoolDesc(ctypes.Structure):
    _fields_ = [("num_levels", ctypes.c_int32), ("batch", ctypes.c_int32), ("rois_per_image", ctypes.c_int32),
                ("channels", ctypes.c_int32), ("pooled_size", ctypes.c_int32), ("sampling_ratio", ctypes.c_int32),
                ("height", ctypes.c_int32 * OSD_MAX_LEVELS), ("width", ctypes.c_int32 * OSD_MAX_LEVELS),
                ("spatial_scale", ctypes.c_float * OSD_MAX_LEVELS),
                ("k_min", ctypes.c_int32), ("k_max", ctypes.c_int32), ("canonical_scale", ctypes.c_float),
                ("canonical_level", ctypes.c_int32), ("eps", ctypes.c_float),
                ("feat", ctypes.c_void_p * OSD_MAX_LEVELS), ("rois", ctypes.c_void_p), ("roi_count", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("levels_out", ctypes.c_void_p), ("workspace", ctypes.c_void_p),
                ("workspace_bytes", ctypes.c_size_t), ("out_nhwc_bf16", ctypes.c_void_p)]


class BoxHeadDesc(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("rois_per_image", ctypes.c_int32), ("channels", ctypes.c_int32),
                ("pooled_size", ctypes.c_int32), ("mlp_dim", ctypes.c_int32), ("num_classes", ctypes.c_int32),
                ("num_box_out", ctypes.c_int32), ("roi_chunk", ctypes.c_int32), ("gn_eps", ctypes.c_float),
                ("lrelu_slope", ctypes.c_float), ("pooled", ctypes.c_void_p), ("supp", ctypes.c_void_p),
                ("w1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("gn1_w", ctypes.c_void_p), ("gn1_b", ctypes.c_void_p),
                ("w2", ctypes.c_void_p), ("b2", ctypes.c_void_p), ("gn2_w", ctypes.c_void_p), ("gn2_b", ctypes.c_void_p),
                ("w3", ctypes.c_void_p), ("b3", ctypes.c_void_p), ("gn3_w", ctypes.c_void_p), ("gn3_b", ctypes.c_void_p),
                ("w6", ctypes.c_void_p), ("b6", ctypes.c_void_p), ("w7", ctypes.c_void_p), ("b7", ctypes.c_void_p),
                ("wp", ctypes.c_void_p), ("bp", ctypes.c_void_p),
                ("class_logits", ctypes.c_void_p), ("box_regression", ctypes.c_void_p),
                ("pooled_nhwc_bf16", ctypes.c_void_p)]


# every symbol include/osd_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "osd_version": (ctypes.c_int, []),
    "osd_last_error": (ctypes.c_char_p, []),
    "osd_check_device": (ctypes.c_int, []),
    "osd_launch_count": (ctypes.c_int64, []),
    "osd_reset_launch_count": (None, []),
    "osd_timeline_enable": (None, [ctypes.c_int]),
    "osd_timeline_read": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t]),
    "osd_batched_nms_plan": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(NmsPlan)]),
    "osd_batched_nms": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_float,
                                       ctypes.c_int, c_void_p, ctypes.c_size_t, c_void_p, c_void_p, c_void_p]),
    "osd_fcos_postprocess_plan": (ctypes.c_int, [ctypes.POINTER(FcosConfig), ctypes.POINTER(FcosPlan)]),
    "osd_fcos_postprocess": (ctypes.c_int, [ctypes.POINTER(FcosConfig), ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                            ctypes.POINTER(c_void_p), c_void_p, c_void_p, ctypes.c_size_t, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "osd_match_forward": (ctypes.c_int, [ctypes.POINTER(MatchDesc), c_void_p]),
    "osd_fusion_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(FusionDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "osd_fusion_forward": (ctypes.c_int, [ctypes.POINTER(FusionDesc), c_void_p, ctypes.c_size_t, c_void_p]),
    "osd_support_pool": (ctypes.c_int, [ctypes.POINTER(SupportPoolDesc), c_void_p]),
    "osd_roi_pool": (ctypes.c_int, [ctypes.POINTER(RoiPoolDesc), c_void_p]),
    "osd_roi_pool_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(RoiPoolDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "osd_box_head_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(BoxHeadDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "osd_box_head_forward": (ctypes.c_int, [ctypes.POINTER(BoxHeadDesc), c_void_p, ctypes.c_size_t, c_void_p]),
    "osd_coco_records": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int32, ctypes.c_int32,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "osd_coco_write_json": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p, ctypes.c_int32,
                                           ctypes.c_char_p]),
    "osd_comm_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(c_void_p)]),
    "osd_comm_free": (ctypes.c_int, [c_void_p]),
    "osd_comm_export": (ctypes.c_int, [c_void_p, ctypes.c_char_p]),
    "osd_comm_import": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "osd_comm_close": (ctypes.c_int, [c_void_p]),
    "osd_comm_push": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int32, c_void_p, ctypes.c_size_t, c_void_p]),
    "osd_comm_push_kernel": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int32, c_void_p, ctypes.c_size_t, c_void_p]),
    "osd_box_postprocess_plan": (ctypes.c_int, [ctypes.POINTER(BoxPostConfig), ctypes.POINTER(BoxPostPlan)]),
    "osd_box_postprocess": (ctypes.c_int, [ctypes.POINTER(BoxPostConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, ctypes.c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_device_ok = set()


class OsdError(RuntimeError):
    pass


def load():
    """Load libosd_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise OsdError(
                    f"{LIB_PATH} is missing. Build it with `python -m oneshotdet_b200.build` "
                    "(nvcc, sm_100a). oneshotdet_b200 has no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)  # AttributeError if the export is missing
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().osd_last_error().decode("utf-8", "replace")
        raise OsdError(f"{what or 'libosd_b200'} failed (status {rc}): {msg}")


def require_device(device) -> None:
    """The tensors must live on an sm_100 GPU; anything else is an error, never a fallback."""
    import torch

    if device.type != "cuda":
        raise OsdError(f"oneshotdet_b200 runs on B200 (sm_100a) only; got a tensor on '{device}'. There is no CPU path.")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_ok:
        return
    with torch.cuda.device(idx):
        check(load().osd_check_device(), "osd_check_device")
    _device_ok.add(idx)


def current_stream_ptr(device) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


class Workspace:
    """Grow-only scratch buffer per (device, stream) from PyTorch's caching allocator.  Reuse is stream-ordered, which is
    only sound within ONE stream: calls issued on different streams get different buffers, and a buffer that is replaced by
    a larger one goes back to the allocator on the stream that used it."""

    def __init__(self):
        self._buf = {}

    def get(self, device, nbytes: int):
        import torch

        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (device.type, idx, torch.cuda.current_stream(idx).cuda_stream)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf
