"""Observed error of the fusion module against the pure-fp32 oracle at the shapes of tests/test_gpu_fusion.py::test_full_module
(to set that test's bounds at 2x the observed values)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from oneshotdet_b200 import fusion  # noqa: E402

for b, s, c, h, w in [(2, 1, 64, 96, 160), (1, 2, 128, 72, 104), (2, 1, 256, 136, 200)]:
    feats, supp = orc.synth_features(b, s, c, h, w, seed=41 + c)
    module = orc.make_compress_dim_conv(c, seed=6)
    got = fusion.fusion_forward([f.cuda() for f in feats], [x.cuda() for x in supp], b, module, stage="full")
    ref = orc.match_fusion(feats, supp, b, module, stage="full")
    mx = max(float((g.cpu() - r).abs().max()) for g, r in zip(got, ref))
    mean = max(float((g.cpu() - r).abs().mean()) for g, r in zip(got, ref))
    print(f"C={c}: max|err| {mx:.4f}  mean|err| {mean:.5f}  max|ref| {max(float(r.abs().max()) for r in ref):.2f}")
