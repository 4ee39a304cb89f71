"""The matching module: support embedding x target FPN maps.

The reference has no module for this -- the product is inlined at
maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:306-311 (with the K-shot mean of :100-104) and the
concat / 1x1-fusion forms live in the second stage (modeling/roi_heads/box_head/box_head.py:43-54, :144-149).
``MatchingModule.forward(features, supp_pooled, batch_size)`` takes exactly the tensors that loop consumes and
returns the list of correlation features the FCOS head consumes."""
from __future__ import annotations

import torch
from torch import nn

from oneshotdet_b200 import ops


class MatchingModule(nn.Module):
    MODES = ("product", "concat", "concat_reversed", "fusion")

    def __init__(self, mode: str = "product", channels: int = 256):
        super().__init__()
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {self.MODES}")
        self.mode = mode
        self.channels = channels
        if mode == "fusion":
            c = channels
            # parameter names follow roi_heads.box.compress_dim_conv.{0,1,3,4} (box_head.py:43-54) so a reference
            # state_dict loads unchanged
            self.compress_dim_conv = nn.Sequential(
                nn.Conv2d(2 * c, 2 * c, 1), nn.GroupNorm(32, 2 * c), nn.LeakyReLU(0.2),
                nn.Conv2d(2 * c, c, 1), nn.GroupNorm(32, c), nn.LeakyReLU(0.2))
            for layer in self.compress_dim_conv:
                if isinstance(layer, nn.Conv2d):
                    nn.init.normal_(layer.weight, std=0.01)  # box_head.py:52-54

    @torch.no_grad()
    def forward(self, features, supp_pooled, batch_size=None):
        """features: 5 x [B,C,H_l,W_l]; supp_pooled: 5 x [B*S,C,1,1] (episode-major, shot-minor, the output of
        supp_pooling at generalized_rcnn.py:303-305, *before* batch_pooling)."""
        if batch_size is None:
            batch_size = features[0].size(0)
        if self.mode == "fusion":
            from oneshotdet_b200 import fusion  # noqa: PLC0415

            return fusion.fusion_forward(features, supp_pooled, batch_size, self.compress_dim_conv)
        return ops.match_forward(list(features), list(supp_pooled), batch_size, self.mode)
