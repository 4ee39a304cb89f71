"""Mirror of the hot-path functions of maskrcnn_benchmark/structures/boxlist_ops.py: ``boxlist_nms``
(:10-34), ``remove_small_boxes`` (:202-216), ``cat_boxlist`` (:270-298)."""
from __future__ import annotations

import torch

from oneshotdet_b200.layers import nms as _box_nms

from .bounding_box import BoxList


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    """NMS on a BoxList with the scores in ``score_field``; ``nms_thresh <= 0`` returns the input
    unchanged (boxlist_ops.py:22-23); ``max_proposals > 0`` keeps the first ``max_proposals`` survivors in
    ascending index order (boxlist_ops.py:31-32)."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    keep = keep.to(boxlist.bbox.device)  # empty input comes back as a CPU tensor (csrc/nms.h:17-18)
    return boxlist[keep].convert(mode)


def remove_small_boxes(boxlist, min_size):
    """Keep boxes whose width and height (legacy +1) are both >= min_size."""
    wh = boxlist.convert("xywh").bbox
    keep = ((wh[:, 2] >= min_size) & (wh[:, 3] >= min_size)).nonzero().squeeze(1)
    return boxlist[keep]


def cat_boxlist(bboxes):
    """Concatenate BoxLists of one image (same size, mode and fields)."""
    assert isinstance(bboxes, (list, tuple)) and len(bboxes) > 0
    assert all(isinstance(b, BoxList) for b in bboxes)
    size, mode = bboxes[0].size, bboxes[0].mode
    assert all(b.size == size for b in bboxes)
    assert all(b.mode == mode for b in bboxes)
    fields = set(bboxes[0].fields())
    assert all(set(b.fields()) == fields for b in bboxes)
    if len(bboxes) == 1:
        return bboxes[0]
    out = BoxList(torch.cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in fields:
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], dim=0))
    return out
