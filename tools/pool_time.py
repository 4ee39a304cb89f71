"""Times the second stage's ROI pooler at 16 x 2000 ROIs, C = 256 (fp32 [roi,C,7,7] and bf16 [roi,49,C] outputs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oneshotdet_b200 as osd  # noqa: E402

dev = torch.device("cuda:0")
b, r, c, h, w = 16, 2000, 256, 800, 1344
g = torch.Generator(device=dev).manual_seed(4242)
feats = [torch.empty((b, c, -(-h // s), -(-w // s)), device=dev).normal_(generator=g) for s in (8, 16, 32, 64, 128)]
ctr = torch.rand((b, r, 2), device=dev, generator=g) * torch.tensor([1333.0, 800.0], device=dev)
wh = torch.rand((b, r, 2), device=dev, generator=g) * 400.0 + 16.0
rois = torch.cat((ctr - wh / 2, ctr + wh / 2), 2).clamp_(min=0.0).contiguous()
pooler = osd.Pooler((7, 7), [1 / s for s in (8, 16, 32, 64, 128)], 2)
for rep in range(3):
    for rows in (True, False):
        for _ in range(2):
            pooler.forward_fixed(feats, rois, rows_bf16=rows)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            pooler.forward_fixed(feats, rois, rows_bf16=rows)
        ev[1].record()
        torch.cuda.synchronize()
        print("rows_bf16" if rows else "fp32", round(ev[0].elapsed_time(ev[1]) / 5, 3), "ms")
