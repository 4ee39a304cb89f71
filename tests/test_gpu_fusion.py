"""GPU parity: the tcgen05 1x1 fusion conv (csrc/fusion_conv.cu) against the oracle = the reference's
``compress_dim_conv`` nn.Sequential (box_head.py:43-54) on cat((x, support), 1), torch fp32 on the CPU.

Tolerances (north_star: 'bf16 1x1-conv mode within stated tolerance'):
  * against the oracle evaluated on bf16-ROUNDED operands (same products, fp32 accumulation, different summation
    order): conv1 |err| <= 2e-4 + 1e-3*|ref|  -- this is the check that catches layout / descriptor mistakes;
  * against the pure fp32 oracle: conv1 max|err| <= 2 % of max|ref| (bf16 has 8 mantissa bits, K = C terms);
    full module (after two GroupNorms, outputs are O(1)): max|err| <= 0.04, mean|err| <= 0.0035 (2x the observed
    0.019 / 0.0017); the executed-reference fixtures keep the looser 0.08 (tiny levels normalise over few values)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(feats, supp, b, module, stage):
    from oneshotdet_b200 import fusion

    out = fusion.fusion_forward([f.to(DEV) for f in feats], [s.to(DEV) for s in supp], b, module, stage=stage)
    torch.cuda.synchronize()
    return [o.cpu() for o in out]


def bf16_emulated(feats, supp, b, module, stage):
    """Oracle on bf16-rounded operands, fp32 accumulation -- what the tensor core computes."""
    import copy

    m = copy.deepcopy(module)
    c = feats[0].shape[1]
    with torch.no_grad():
        w1 = m[0].weight.clone()
        w1[:, :c] = w1[:, :c].bfloat16().float()      # the support half stays fp32 (it is folded into the bias)
        m[0].weight.copy_(w1)
        m[3].weight.copy_(m[3].weight.bfloat16().float())
        outs = []
        for f, s in zip(feats, supp):
            p = orc.batch_pooling(s, b).expand(-1, -1, f.shape[2], f.shape[3])
            x = torch.cat((f.bfloat16().float(), p), dim=1)
            y1 = m[0](x)
            if stage == "conv1":
                outs.append(y1)
                continue
            a1 = m[2](m[1](y1)).bfloat16().float()
            outs.append(m[5](m[4](m[3](a1))))
    return outs


@pytest.mark.parametrize("b,s,c,h,w", [(1, 1, 64, 64, 64), (2, 1, 64, 96, 160), (2, 3, 128, 72, 104), (1, 1, 256, 128, 192),
                                       (2, 2, 256, 200, 264)])
def test_conv1_stage(b, s, c, h, w):
    feats, supp = orc.synth_features(b, s, c, h, w, seed=31 + c)
    module = orc.make_compress_dim_conv(c, seed=5)
    got = run(feats, supp, b, module, "conv1")
    emu = bf16_emulated(feats, supp, b, module, "conv1")
    ref = orc.match_fusion(feats, supp, b, module, stage="conv1")
    for g, e, r in zip(got, emu, ref):
        assert g.shape == r.shape
        err = (g - e).abs()
        assert bool((err <= 2e-4 + 1e-3 * e.abs()).all()), f"max err vs bf16-emulated oracle {err.max():.3e}"
        assert float((g - r).abs().max()) <= 0.02 * float(r.abs().max())


@pytest.mark.parametrize("b,s,c,h,w", [(2, 1, 64, 96, 160), (1, 2, 128, 72, 104), (2, 1, 256, 136, 200)])
def test_full_module(b, s, c, h, w):
    feats, supp = orc.synth_features(b, s, c, h, w, seed=41 + c)
    module = orc.make_compress_dim_conv(c, seed=6)
    got = run(feats, supp, b, module, "full")
    emu = bf16_emulated(feats, supp, b, module, "full")
    ref = orc.match_fusion(feats, supp, b, module, stage="full")
    for g, e, r in zip(got, emu, ref):
        assert g.shape == r.shape
        assert float((g - e).abs().max()) <= 5e-3, f"vs bf16-emulated oracle: {float((g - e).abs().max()):.3e}"
        d = (g - r).abs()
        # observed on B200 (tools/fusion_err.py): max 0.016-0.019, mean 0.0015-0.0017 on outputs of magnitude <= 4.6;
        # the bounds are twice that
        assert float(d.max()) <= 0.04 and float(d.mean()) <= 0.0035, (float(d.max()), float(d.mean()))


def test_full_module_baseline_geometry():
    """BASELINE configs[1] geometry: all five levels of an 800x1344 target, C = 256, 2 episodes (partial 128-pixel tiles
    on every level but P3, rows that are not 16-byte multiples on P5-P7)."""
    b, c = 2, 256
    feats, supp = orc.synth_features(b, 1, c, 800, 1344, seed=97)
    module = orc.make_compress_dim_conv(c, seed=8)
    got = run(feats, supp, b, module, "full")
    emu = bf16_emulated(feats, supp, b, module, "full")
    for l, (g, e) in enumerate(zip(got, emu)):
        assert g.shape == e.shape
        # tiny levels (7x11) normalise over few values: a rounding difference moves the statistics more there
        lim = 5e-3 if e[0, 0].numel() >= 256 else 2e-2
        assert float((g - e).abs().max()) <= lim, (l, float((g - e).abs().max()))


@pytest.mark.parametrize("name", ["s1_c64", "s3_c64"])
def test_reference_fixtures(golden_dir, name):
    """Outputs of the reference's own compress_dim_conv executed in the build container."""
    z = np.load(os.path.join(golden_dir, f"match_{name}.npz"))
    b, c = int(z["batch"]), int(z["channels"])
    nl = len([k for k in z.files if k.startswith("feat")])
    feats = [torch.from_numpy(z[f"feat{l}"]) for l in range(nl)]
    supp = [torch.from_numpy(z[f"supp{l}"]) for l in range(nl)]
    module = orc.make_compress_dim_conv(c)
    module.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")})
    conv1 = run(feats, supp, b, module, "conv1")
    full = run(feats, supp, b, module, "full")
    for l in range(nl):
        r1 = torch.from_numpy(z[f"conv1_{l}"])
        assert float((conv1[l] - r1).abs().max()) <= 0.02 * float(r1.abs().max())
        rf = torch.from_numpy(z[f"fused{l}"])
        d = (full[l] - rf).abs()
        # tiny levels (2x3 pixels) have GroupNorm statistics over a handful of values: looser bound there
        lim = 0.08 if rf[0, 0].numel() >= 64 else 0.25
        assert float(d.max()) <= lim, (l, float(d.max()))


def test_matching_module_fusion_mode_loads_reference_state_dict():
    import oneshotdet_b200 as osd

    c = 64
    ref_module = orc.make_compress_dim_conv(c, seed=9)
    m = osd.MatchingModule("fusion", channels=c)
    m.compress_dim_conv.load_state_dict(ref_module.state_dict())   # same parameter names as box_head.py:43-54
    feats, supp = orc.synth_features(2, 1, c, 64, 96, seed=3)
    out = m.to(DEV)([f.to(DEV) for f in feats], [s.to(DEV) for s in supp], 2)
    ref = orc.match_fusion(feats, supp, 2, ref_module)
    for o, r in zip(out, ref):
        assert float((o.cpu() - r).abs().max()) <= 0.08


@pytest.mark.parametrize("mode", ["single", "mc", "pair", "gram"])
def test_cluster_modes_in_subprocess(mode):
    """The fused kernels' other launch forms (the switches are read once per process): clusters of eight CTAs with
    multicast weight stages ('mc'), CTA pairs on cta_group::2 ('pair'), the plain one-CTA form, and pass A taking the
    GroupNorm-1 statistics from a Gram GEMM ('gram', OSD_FUSION_GRAM=1) must all pass the full-module parity tests of
    this file."""
    import subprocess
    import sys

    if os.environ.get("OSD_FUSION_SUBTEST"):
        pytest.skip("already inside the subprocess")
    env = dict(os.environ, OSD_FUSION_SUBTEST="1")
    if mode == "gram":
        env.update(OSD_FUSION_MODE="single", OSD_FUSION_GRAM="1")
    else:
        env.update(OSD_FUSION_MODE=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "full_module or reference_fixtures", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
