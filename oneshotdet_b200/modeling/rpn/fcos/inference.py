"""Drop-in for ``maskrcnn_benchmark.modeling.rpn.fcos.inference.FCOSPostProcessor`` (inference.py:18-323)
and ``make_fcos_postprocessor`` (:325-364), backed by the fused sm_100a pipeline
(score -> per-level top-k -> decode/clip -> batched NMS -> post-NMS top-n) of libosd_b200.so.

Same constructor arguments, same ``forward`` signature, same return contract: one ``BoxList`` per episode,
``size = (w, h)``, mode xyxy, a single ``scores`` field; rows ordered by descending score when more than
``fpn_post_nms_top_n`` boxes survive NMS, otherwise in ascending candidate order (levels P3->P7, locations
row-major).  One host synchronisation per call (to size the BoxLists) instead of two ``.item()`` per image and
level; ``forward_fixed`` returns the padded device tensors with no synchronisation at all."""
from __future__ import annotations

import torch

from oneshotdet_b200 import ops
from oneshotdet_b200.structures.bounding_box import BoxList

DEFAULT_FPN_STRIDES = (8, 16, 32, 64, 128)  # config/defaults.py:299


class FCOSPostProcessor(torch.nn.Module):
    def __init__(self, config, pre_nms_thresh, pre_nms_top_n, nms_thresh, fpn_post_nms_top_n, min_size, num_classes,
                 dense_points, score_calculator, fpn_strides=None, strict_iou=False, early_exit=True):
        super().__init__()
        self.cfg = config
        self.pre_nms_thresh = pre_nms_thresh
        self.pre_nms_top_n = pre_nms_top_n
        self.nms_thresh = nms_thresh
        self.fpn_post_nms_top_n = fpn_post_nms_top_n
        self.min_size = min_size
        self.num_classes = num_classes
        self.dense_points = dense_points
        self.score_calculator = score_calculator
        if fpn_strides is None:
            try:
                fpn_strides = tuple(config.MODEL.FCOS.FPN_STRIDES)
            except AttributeError:
                fpn_strides = DEFAULT_FPN_STRIDES
        self.fpn_strides = tuple(int(s) for s in fpn_strides)
        # False: IoU >= thr suppresses, as the runnable reference nms_cpu (csrc/cpu/nms_cpu.cpp:60);
        # True: IoU > thr, as the reference CUDA kernel (csrc/cuda/nms.cu:60)
        self.strict_iou = strict_iou
        self.early_exit = early_exit
        self._checked_grids = set()
        if score_calculator != "BINARY":
            if score_calculator == "MULTI":
                raise NotImplementedError("score_calculator='MULTI' (inference.py:63-65) is not on the accelerated path; "
                                          "the shipped configs use LOSS.CLS_LOSS='BINARY' (config/defaults.py:549)")
            raise Exception("loss type wrong")  # inference.py:67
        if dense_points != 1:
            raise NotImplementedError("dense_points > 1 (fcos.py:236-248) is not supported; MODEL.FCOS.DENSE_POINTS defaults to 1")

    # -- helpers ---------------------------------------------------------------------------------
    def _check_locations(self, locations, box_cls):
        """The kernel recomputes the (x, y) grid of fcos.py:220-234 from the stride; verify once per grid shape that
        the ``locations`` the caller passes are that grid."""
        for loc, c, s in zip(locations, box_cls, self.fpn_strides):
            h, w = c.shape[-2:]
            key = (h, w, s, loc.device)
            if key in self._checked_grids:
                continue
            ys, xs = torch.meshgrid(torch.arange(h, device=loc.device), torch.arange(w, device=loc.device), indexing="ij")
            grid = torch.stack((xs.reshape(-1), ys.reshape(-1)), dim=1).to(torch.float32) * s + s // 2
            if tuple(loc.shape) != tuple(grid.shape) or not torch.equal(loc.to(torch.float32), grid):
                raise ValueError(f"locations of the {h}x{w} level are not the stride-{s} FCOS grid (fcos.py:220-234)")
            self._checked_grids.add(key)

    def _pos_logits(self, box_cls):
        # inference.py:57-61: with extra (negative-support) classes only channel 0 is the positive class
        return [c if c.size(1) == 1 else c[:, 0:1].contiguous() for c in box_cls]

    # -- the accelerated path ----------------------------------------------------------------------
    def forward_fixed(self, box_cls, box_regression, centerness, image_sizes, reg_scales=None) -> ops.FcosResult:
        """Padded device-resident result (boxes [B,K,4], scores [B,K], index [B,K], count [B]); no host sync.

        ``reg_scales`` (one float per level, the head's ``scales[l].scale``): ``box_regression`` then holds the RAW
        ``bbox_pred`` conv outputs and the head's tail ``torch.exp(self.scales[l](x))`` (fcos.py:95-97) is folded
        into the decode of the selected locations."""
        if len(box_cls) > len(self.fpn_strides):
            raise ValueError("more feature levels than fpn_strides")
        return ops.fcos_postprocess(self._pos_logits(box_cls), box_regression, centerness,
                                    self.fpn_strides[:len(box_cls)], image_sizes, self.pre_nms_thresh,
                                    self.pre_nms_top_n, self.nms_thresh, self.fpn_post_nms_top_n, self.min_size,
                                    strict=self.strict_iou, early_exit=self.early_exit, reg_scales=reg_scales)

    def forward(self, locations, box_cls, box_regression, centerness, image_sizes, targets=None):
        """inference.py:251-281 at eval.  ``locations`` is accepted for interface compatibility (and verified);
        ``image_sizes`` is a list of (h, w)."""
        if self.training and targets is not None and not self.cfg.MODEL.RPN_ONLY:
            raise NotImplementedError("training-time proposal augmentation (inference.py:139-249, :273-279) is outside "
                                      "the inference hot path; call the reference implementation for training")
        if locations is not None:
            self._check_locations(locations, box_cls)
        res = self.forward_fixed(box_cls, box_regression, centerness, image_sizes)
        counts = res.count.tolist()  # the one host sync
        boxlists = []
        for i, n in enumerate(counts):
            h, w = image_sizes[i]
            bl = BoxList(res.boxes[i, :n], (int(w), int(h)), mode="xyxy")
            bl.add_field("scores", res.scores[i, :n])
            boxlists.append(bl)
        return boxlists


def make_fcos_postprocessor(config, is_train):
    """inference.py:325-364: single-stage (RPN_ONLY) values come from MODEL.FCOS / TEST, two-stage values from
    MODEL.RPN."""
    if config.MODEL.RPN_ONLY:
        pre_nms_thresh = config.MODEL.FCOS.INFERENCE_TH
        pre_nms_top_n = config.MODEL.FCOS.PRE_NMS_TOP_N
        nms_thresh = config.MODEL.FCOS.NMS_TH
        fpn_post_nms_top_n = config.TEST.DETECTIONS_PER_IMG
        dense_points = config.MODEL.FCOS.DENSE_POINTS
        num_cls = config.MODEL.FCOS.NUM_CLASSES
        if config.FEW_SHOT.NEG_SUPPORT.TURN_ON:
            num_cls += config.FEW_SHOT.NEG_SUPPORT.NUM_CLS
        score_calculator = config.LOSS.CLS_LOSS
        min_size = 0
    else:
        num_cls = 2
        pre_nms_thresh = 0
        fpn_post_nms_top_n = config.MODEL.RPN.FPN_POST_NMS_TOP_N_TEST if not is_train \
            else config.MODEL.RPN.FPN_POST_NMS_TOP_N_TRAIN
        dense_points = config.MODEL.FCOS.DENSE_POINTS
        pre_nms_top_n = config.MODEL.RPN.PRE_NMS_TOP_N_TEST if not is_train else config.MODEL.RPN.PRE_NMS_TOP_N_TRAIN
        nms_thresh = config.MODEL.RPN.NMS_THRESH
        min_size = config.MODEL.RPN.MIN_SIZE
        score_calculator = config.LOSS.CLS_LOSS
    return FCOSPostProcessor(config=config, pre_nms_thresh=pre_nms_thresh, pre_nms_top_n=pre_nms_top_n,
                             nms_thresh=nms_thresh, fpn_post_nms_top_n=fpn_post_nms_top_n, min_size=min_size,
                             num_classes=num_cls, dense_points=dense_points, score_calculator=score_calculator)
