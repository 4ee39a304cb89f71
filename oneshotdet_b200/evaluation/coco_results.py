"""Result hand-off: fixed-shape detections -> COCO detection records -> the JSON file the
reference writes in ``prepare_for_coco_detection`` (maskrcnn_benchmark/data/datasets/evaluation/coco/coco_eval.py:70-176).

``coco_records`` is the device half (one kernel: resize to the original image size, xyxy -> xywh, compaction);
``write_coco_json`` the host half (native formatter, byte-identical to ``json.dump(..., sort_keys=True, indent=4,
separators=(',', ':'))``); ``prepare_for_coco_detection`` strings them together for a list of ``BoxList`` the way the
reference's loop body does (:137-156).  The dataset bookkeeping around it (pycocotools ground truth, id maps) is the
caller's."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import OsdError


@torch.no_grad()
def coco_records(boxes, scores, count, det_sizes_wh, orig_sizes_wh):
    """boxes [E,K,4], scores [E,K], count int32 [E] (CUDA, as FcosResult / BoxPostResult hold them); det_sizes_wh /
    orig_sizes_wh: [E,2] (w, h) lists or int32 tensors.  Returns (records [n,5] = x, y, w, h, score; episode [n] int32),
    both on the device; ``n`` is read back (one host sync)."""
    lib = _lib.load()
    dev = boxes.device
    _lib.require_device(dev)
    e, k = boxes.size(0), boxes.size(1)
    if boxes.dtype != torch.float32 or scores.dtype != torch.float32 or count.dtype != torch.int32:
        raise OsdError("coco_records: boxes / scores must be float32 and count int32")
    if tuple(scores.shape) != (e, k) or count.numel() != e:
        raise OsdError("coco_records: shapes of boxes, scores and count disagree")

    def as_wh(x):
        t = x if isinstance(x, torch.Tensor) else torch.tensor([[int(a), int(b)] for a, b in x], dtype=torch.int32)
        t = t.to(device=dev, dtype=torch.int32).contiguous()
        if t.numel() != 2 * e:
            raise OsdError(f"coco_records: {e} episodes but {t.numel() // 2} sizes")
        return t

    det, orig = as_wh(det_sizes_wh), as_wh(orig_sizes_wh)
    rec = torch.empty((e * k, 5), dtype=torch.float32, device=dev)
    rec_ep = torch.empty((e * k,), dtype=torch.int32, device=dev)
    total = torch.zeros((1,), dtype=torch.int32, device=dev)
    b, s, c = boxes.contiguous(), scores.contiguous(), count.contiguous()
    with torch.cuda.device(dev):
        rc = lib.osd_coco_records(b.data_ptr(), s.data_ptr(), c.data_ptr(), det.data_ptr(), orig.data_ptr(), e, k,
                                  rec.data_ptr(), rec_ep.data_ptr(), total.data_ptr(), _lib.current_stream_ptr(dev))
    _lib.check(rc, "osd_coco_records")
    n = int(total.item())
    return rec[:n], rec_ep[:n]


def write_coco_json(records, record_episode, image_ids, category_ids, path: str) -> None:
    """records [n,5] (x, y, w, h, score), record_episode [n]: tensors (any device) or numpy arrays; image_ids /
    category_ids: one integer per episode.  Writes the reference's ``coco_custom_result.json`` byte for byte."""
    lib = _lib.load()

    def host(a, dtype):
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(a, dtype=dtype)

    rec = host(records, np.float32).reshape(-1, 5)
    ep = host(record_episode, np.int32).reshape(-1)
    img = host(np.asarray(image_ids), np.int64).reshape(-1)
    cat = host(np.asarray(category_ids), np.int64).reshape(-1)
    if ep.shape[0] != rec.shape[0] or img.shape[0] != cat.shape[0]:
        raise OsdError("write_coco_json: inconsistent lengths")
    rc = lib.osd_coco_write_json(rec.ctypes.data_as(ctypes.c_void_p), ep.ctypes.data_as(ctypes.c_void_p), rec.shape[0],
                                 img.ctypes.data_as(ctypes.c_void_p), cat.ctypes.data_as(ctypes.c_void_p), img.shape[0],
                                 str(path).encode())
    _lib.check(rc, "osd_coco_write_json")


def prepare_for_coco_detection(predictions, img_infos, category_ids, path="coco_custom_result.json", image_ids=None):
    """The detection part of the reference function for a list of ``BoxList`` (field ``scores``): predictions[i] is
    resized to (img_infos[i]['width'], img_infos[i]['height']), converted to xywh and written with image_id i
    (coco_eval.py:137-156: ``image_id`` is the running index, ``category_id`` the episode's class)."""
    e = len(predictions)
    if image_ids is None:
        image_ids = list(range(e))
    if e == 0:
        write_coco_json(np.zeros((0, 5), np.float32), np.zeros((0,), np.int32), [], [], path)
        return 0
    dev = predictions[0].bbox.device
    k = max(1, max(len(p) for p in predictions))
    boxes = torch.zeros((e, k, 4), dtype=torch.float32, device=dev)
    scores = torch.zeros((e, k), dtype=torch.float32, device=dev)
    for i, p in enumerate(predictions):
        n = len(p)
        if n:
            boxes[i, :n] = p.convert("xyxy").bbox
            scores[i, :n] = p.get_field("scores")
    count = torch.tensor([len(p) for p in predictions], dtype=torch.int32, device=dev)
    det = [tuple(p.size) for p in predictions]
    orig = [(int(info["width"]), int(info["height"])) for info in img_infos]
    rec, ep = coco_records(boxes, scores, count, det, orig)
    write_coco_json(rec, ep, image_ids, category_ids, path)
    return rec.size(0)
