"""GPU parity: the matching stream kernels against the oracle (= the reference expressions) and the
executed-reference fixtures.  fp32 product/concat are bit-exact (one rounded multiply per element, K-shot mean =
sequential sum / S exactly as ATen computes it); north_star's bound is 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(feats, supp, b, mode, channels_last=False):
    import oneshotdet_b200 as osd

    f = [x.to(DEV) for x in feats]
    if channels_last:
        f = [x.contiguous(memory_format=torch.channels_last) for x in f]
    out = osd.match_forward(f, [s.to(DEV) for s in supp], b, mode)
    if channels_last:
        assert all(o.is_contiguous(memory_format=torch.channels_last) for o in out)
    return [o.cpu() for o in out]


@pytest.mark.parametrize("name", ["s1_c64", "s3_c64"])
def test_reference_fixtures(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"match_{name}.npz"))
    b, c = int(z["batch"]), int(z["channels"])
    nl = len([k for k in z.files if k.startswith("feat")])
    feats = [torch.from_numpy(z[f"feat{l}"]) for l in range(nl)]
    supp = [torch.from_numpy(z[f"supp{l}"]) for l in range(nl)]
    prod = run(feats, supp, b, "product")
    cat = run(feats, supp, b, "concat")
    for l in range(nl):
        np.testing.assert_array_equal(prod[l].numpy(), z[f"product{l}"])
        np.testing.assert_array_equal(cat[l][:, :c].numpy(), z[f"feat{l}"])
        exp = np.broadcast_to(z[f"pooled{l}"], z[f"feat{l}"].shape)
        np.testing.assert_array_equal(cat[l][:, c:].numpy(), exp)


@pytest.mark.parametrize("b,s,c,h,w", [(2, 1, 256, 200, 336), (3, 5, 256, 104, 104), (1, 2, 32, 72, 40),
                                       (2, 3, 20, 37, 53), (1, 1, 4, 8, 8), (5, 1, 3, 17, 9)])
@pytest.mark.parametrize("mode", ["product", "concat", "concat_reversed"])
def test_nchw_fp32_bit_exact(b, s, c, h, w, mode):
    if mode != "product" and c % 4:
        pytest.skip("concat needs C % 4 == 0")
    feats, supp = orc.synth_features(b, s, c, h, w, seed=b * 100 + s)
    got = run(feats, supp, b, mode)
    exp = orc.match_product(feats, supp, b) if mode == "product" else \
        orc.match_concat(feats, supp, b, reverse=(mode == "concat_reversed"))
    for g, e in zip(got, exp):
        assert g.shape == e.shape
        assert torch.equal(g, e)


@pytest.mark.parametrize("mode", ["product", "concat", "concat_reversed"])
@pytest.mark.parametrize("s", [1, 5])
def test_channels_last_fp32_bit_exact(mode, s):
    feats, supp = orc.synth_features(2, s, 64, 104, 136, seed=77 + s)
    got = run(feats, supp, 2, mode, channels_last=True)
    exp = orc.match_product(feats, supp, 2) if mode == "product" else \
        orc.match_concat(feats, supp, 2, reverse=(mode == "concat_reversed"))
    for g, e in zip(got, exp):
        assert torch.equal(g, e)


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("s", [1, 3])
def test_bf16_io(channels_last, s):
    """bf16 in / bf16 out, fp32 multiply: equals torch's own bf16 arithmetic (float multiply, one rounding)."""
    feats, supp = orc.synth_features(2, s, 64, 72, 104, seed=5 + s)
    feats = [f.bfloat16() for f in feats]
    supp = [x.bfloat16() for x in supp]
    got = run(feats, supp, 2, "product", channels_last)
    exp = orc.match_product(feats, supp, 2)
    for g, e in zip(got, exp):
        assert g.dtype == torch.bfloat16
        assert torch.equal(g, e)
    gc = run(feats, supp, 2, "concat", channels_last)
    ec = orc.match_concat(feats, supp, 2)
    for g, e in zip(gc, ec):
        assert torch.equal(g, e)


def test_full_size_linearity_and_identity():
    """BASELINE geometry (800x1344, C=256, 2 episodes): size-independent properties instead of an oracle pass --
    a support of ones is the identity, and the product is linear in the support."""
    import oneshotdet_b200 as osd

    g = torch.Generator(device="cpu").manual_seed(1)
    shapes = orc.level_shapes(800, 1344)
    feats = [torch.randn(2, 256, h, w, generator=g).to(DEV) for h, w in shapes]
    ones = [torch.ones(2, 256, 1, 1, device=DEV) for _ in shapes]
    out = osd.match_forward(feats, ones, 2, "product")
    for o, f in zip(out, feats):
        assert torch.equal(o, f)
    supp = [torch.randn(2, 256, 1, 1, generator=g).to(DEV) for _ in shapes]
    a = osd.match_forward(feats, supp, 2, "product")
    b2 = osd.match_forward(feats, [2.0 * s for s in supp], 2, "product")
    for x, y, f, s in zip(a, b2, feats, supp):
        assert torch.equal(2.0 * x, y)            # exact: scaling by 2 commutes with rounding
        assert torch.equal(x, f * s)              # the reference expression itself, on the device


def test_matching_module_dropin():
    import oneshotdet_b200 as osd

    feats, supp = orc.synth_features(2, 2, 32, 64, 64, seed=3)
    m = osd.MatchingModule("product", channels=32)
    out = m([f.to(DEV) for f in feats], [s.to(DEV) for s in supp], 2)
    for o, e in zip(out, orc.match_product(feats, supp, 2)):
        assert torch.equal(o.cpu(), e)


def test_static_and_dynamic_chunk_scheduling_agree(monkeypatch):
    """The bulk-copy kernel draws chunks from a global counter (default) or, when the dedicated counter slots of captured
    launches run out, round-robin; both must give the reference product, also for many back-to-back launches (the
    counters are reset by the last CTA of each launch)."""
    import oneshotdet_b200 as osd

    feats, supp = orc.synth_features(2, 2, 64, 200, 264, seed=77)
    want = orc.match_product(feats, supp, 2)
    df, ds = [f.to(DEV) for f in feats], [x.to(DEV) for x in supp]
    for mode in ("dynamic", "static"):
        if mode == "static":
            monkeypatch.setenv("OSD_MATCH_SCHED", "static")
        for _ in range(70):                       # more launches than counter slots
            out = osd.match_forward(df, ds, 2)
        torch.cuda.synchronize()
        for o, e in zip(out, want):
            assert torch.equal(o.cpu(), e), mode
