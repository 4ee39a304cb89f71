"""Drop-in for the second-stage ``PostProcessor`` (maskrcnn_benchmark/modeling/roi_heads/box_head/inference.py:13-167):
same constructor arguments, same ``forward(x, boxes, cyclic=False, target_ids=None)`` contract, one fused
C-ABI call (``osd_box_postprocess``) for all images instead of the per-image Python loop.  SURVEY 8(f) row 2.

What the reference does per image (:96-104, :133-167) and this module reproduces:
class probability (:62-70) -> ``BoxCoder.decode`` of regression columns [4, 8) against the proposals (:80-82,
box_coder.py:52-95) -> ``clip_to_image`` (:102) -> ``scores[:, 1] > score_thresh`` (:142-146) -> ``boxlist_nms`` (:150) ->
field ``labels = target_id`` (:155-158) -> detections-per-image cut in score order (:162-166)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .... import ops
from ....structures.bounding_box import BoxList


class BoxCoder:
    """Holder of the decode parameters (modeling/box_coder.py:13-21); the arithmetic of ``decode`` (:52-95) runs inside
    the fused kernel."""

    def __init__(self, weights, bbox_xform_clip=math.log(1000. / 16)):
        self.weights = tuple(float(w) for w in weights)
        self.bbox_xform_clip = float(bbox_xform_clip)


_SCORE_MODE = {"focal_loss": "sigmoid", "ce_loss": "softmax", "cxe_loss": "softmax", "mse_loss": "sigmoid",
               "l1_loss": "sigmoid"}


class PostProcessor(nn.Module):
    def __init__(self, cfg, score_thresh=0.05, nms=0.5, detections_per_img=10000, box_coder=None,
                 cls_agnostic_bbox_reg=False, strict_iou=False):
        super().__init__()
        self.cfg = cfg
        self.score_thresh = score_thresh
        self.nms = nms
        self.detections_per_img = detections_per_img
        self.box_coder = box_coder if box_coder is not None else BoxCoder(weights=(10., 10., 5., 5.))
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.strict_iou = strict_iou

    def _score_mode(self, class_logits):
        loss = self.cfg.FEW_SHOT.SECOND_STAGE_CLS_LOSS
        if loss not in _SCORE_MODE:
            raise ValueError(f"unknown FEW_SHOT.SECOND_STAGE_CLS_LOSS '{loss}'")
        if loss in ("mse_loss", "l1_loss") and class_logits.size(1) != 1:
            # :68-70 concatenates [1 - s, s]; column 1 is the foreground probability only for a single logit
            raise NotImplementedError("mse_loss / l1_loss post-processing expects a single class logit")
        return _SCORE_MODE[loss]

    def forward_fixed(self, x, proposals, image_sizes, roi_count=None) -> ops.BoxPostResult:
        """Device-resident form: proposals [B,R,4] (e.g. the FCOS stage's padded output with its counts), result padded
        to K rows per image; no host synchronisation."""
        class_logits, box_regression = x
        if box_regression.size(1) < 8:
            raise ValueError("box_regression needs the 8 columns the reference slices at inference.py:60")
        # :60 keeps the first 8 columns; class 1 reads [4, 8) of them, with or without cls_agnostic_bbox_reg (:77-78, :147)
        return ops.box_postprocess(class_logits, box_regression, proposals, image_sizes, self.score_thresh, self.nms,
                                   self.detections_per_img, self.box_coder.weights, self._score_mode(class_logits), 4,
                                   roi_count, self.box_coder.bbox_xform_clip, strict=self.strict_iou)

    def forward(self, x, boxes, cyclic=False, target_ids=None):
        """inference.py:46-104: ``boxes`` is a list of BoxList (one per image, equal lengths as the Pooler requires,
        poolers.py:79); returns one BoxList per image with fields ``scores`` and ``labels``."""
        if cyclic:
            raise NotImplementedError("cyclic evaluation (:92-95) is not part of the accelerated path")
        if target_ids is None:
            raise ValueError("target_ids is required at inference (inference.py:97-98)")
        counts = {len(b) for b in boxes}
        if len(counts) != 1:
            raise ValueError(f"all images must carry the same number of proposals, got {sorted(counts)}")
        proposals = torch.stack([b.convert("xyxy").bbox for b in boxes], dim=0)
        sizes = [(b.size[1], b.size[0]) for b in boxes]     # BoxList.size is (w, h)
        res = self.forward_fixed(x, proposals, sizes)
        n_out = res.count.tolist()                          # the one host sync
        out = []
        for i, (n, tid) in enumerate(zip(n_out, target_ids)):
            bl = BoxList(res.boxes[i, :n], boxes[i].size, mode="xyxy")
            bl.add_field("scores", res.scores[i, :n])
            bl.add_field("labels", torch.full((n,), int(tid), dtype=torch.int64, device=res.boxes.device))
            out.append(bl)
        return out


def make_roi_box_post_processor(cfg):
    """inference.py:169-189."""
    box_coder = BoxCoder(weights=cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS)
    return PostProcessor(cfg, cfg.MODEL.ROI_HEADS.SCORE_THRESH, cfg.MODEL.ROI_HEADS.NMS,
                         cfg.MODEL.ROI_HEADS.DETECTIONS_PER_IMG, box_coder, cfg.MODEL.CLS_AGNOSTIC_BBOX_REG)
