"""CPU: pins the python restatement (oracle/oracle.py) of FCOS post-processing and matching to the
fixtures produced by executing the unmodified reference (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import canon

CASES = ["two_stage_small", "stress_small", "minsize_small", "nonms_small"]


def load_fcos(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"fcos_post_{name}.npz"))
    nl = len([k for k in z.files if k.startswith("cls")])
    cls = [torch.from_numpy(z[f"cls{l}"]) for l in range(nl)]
    reg = [torch.from_numpy(z[f"reg{l}"]) for l in range(nl)]
    ctr = [torch.from_numpy(z[f"ctr{l}"]) for l in range(nl)]
    pr = z["params"]
    p = orc.PostParams(float(pr[0]), int(pr[1]), float(pr[2]), int(pr[3]), float(pr[4]))
    sizes = [tuple(int(v) for v in s) for s in z["image_sizes"]]
    outs = [(z[f"out_boxes{i}"], z[f"out_scores{i}"]) for i in range(int(z["batch"]))]
    return z, cls, reg, ctr, p, sizes, outs


def test_fixtures_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "fcos_post_*.npz"))) >= 4
    assert len(glob.glob(os.path.join(golden_dir, "match_*.npz"))) >= 2


@pytest.mark.parametrize("name", CASES)
def test_locations_match_reference(golden_dir, name):
    z, cls, *_ = load_fcos(golden_dir, name)
    for l, (c, s) in enumerate(zip(cls, orc.FPN_STRIDES)):
        h, w = c.shape[-2:]
        np.testing.assert_array_equal(orc.compute_locations_per_level(h, w, s), z[f"loc{l}"])


@pytest.mark.parametrize("name", CASES)
def test_postprocess_matches_reference(golden_dir, name):
    z, cls, reg, ctr, p, sizes, outs = load_fcos(golden_dir, name)
    res = orc.fcos_postprocess(cls, reg, ctr, orc.FPN_STRIDES, sizes, p)
    assert len(res) == len(outs)
    for i, (r, (rb, rs)) in enumerate(zip(res, outs)):
        assert r["boxes"].shape == rb.shape, (name, i)
        if rb.shape[0] == p.fpn_post_nms_top_n:
            # score-descending branch (inference.py:316-321): order is part of the contract
            np.testing.assert_array_equal(r["scores"], rs)
            np.testing.assert_array_equal(r["boxes"], rb)
        else:
            # ascending-candidate-index branch: the reference's candidate order inside a level is
            # whatever topk(sorted=False) returned, so compare as sets
            gb, gs = canon(r["boxes"], r["scores"])
            eb, es = canon(rb, rs)
            np.testing.assert_array_equal(gs, es)
            np.testing.assert_array_equal(gb, eb)
        assert tuple(z[f"out_size{i}"]) == (sizes[i][1], sizes[i][0])  # BoxList.size = (w, h)
        assert list(z[f"out_fields{i}"]) == ["scores"]


def test_candidate_counts_two_stage(golden_dir):
    z, cls, reg, ctr, p, sizes, _ = load_fcos(golden_dir, "two_stage_small")
    cands = orc.fcos_candidates(cls, reg, ctr, orc.FPN_STRIDES, sizes, p)
    per_level_cap = [min(c.shape[-1] * c.shape[-2], p.pre_nms_top_n) for c in cls]
    for cb, cs, lv, lc in cands:
        assert cb.shape[0] == sum(per_level_cap)
        assert np.all(np.diff(lv) >= 0)
        for l in range(len(cls)):
            assert np.all(np.diff(lc[lv == l]) > 0)  # location order inside a level


@pytest.mark.parametrize("name", ["s1_c64", "s3_c64"])
def test_matching_matches_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"match_{name}.npz"))
    b, s, c = int(z["batch"]), int(z["shots"]), int(z["channels"])
    nl = len([k for k in z.files if k.startswith("feat")])
    feats = [torch.from_numpy(z[f"feat{l}"]) for l in range(nl)]
    supp = [torch.from_numpy(z[f"supp{l}"]) for l in range(nl)]
    prod = orc.match_product(feats, supp, b)
    cat = orc.match_concat(feats, supp, b)
    mod = orc.make_compress_dim_conv(c)
    mod.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")})
    conv1 = orc.match_fusion(feats, supp, b, mod, stage="conv1")
    fused = orc.match_fusion(feats, supp, b, mod)
    for l in range(nl):
        np.testing.assert_array_equal(prod[l].numpy(), z[f"product{l}"])
        np.testing.assert_array_equal(cat[l][:, :c].numpy(), z[f"feat{l}"])
        np.testing.assert_array_equal(cat[l][:, c:, 0, 0].numpy(), z[f"pooled{l}"][:, :, 0, 0])
        assert cat[l].shape[1] == 2 * c and s >= 1
        np.testing.assert_allclose(conv1[l].numpy(), z[f"conv1_{l}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(fused[l].numpy(), z[f"fused{l}"], rtol=1e-4, atol=1e-5)


def test_head_tail_restatement_matches_executed_reference(golden_dir):
    """oracle.fcos_head_tail (exp(x * scale), fcos.py:95-97) against the reference's Scale module + torch.exp executed
    in the build container: bit-exact (same ATen ops)."""
    z = np.load(os.path.join(golden_dir, "fcos_head_tail.npz"))
    for l in range(3):
        got = orc.fcos_head_tail(torch.from_numpy(z[f"raw{l}"]), float(z[f"scale{l}"]))
        np.testing.assert_array_equal(got.numpy(), z[f"out{l}"])
