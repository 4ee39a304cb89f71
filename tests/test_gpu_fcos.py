"""GPU parity: the fused FCOS post-processing (score / top-k / decode / clip / NMS / post-top-n) against the
oracle and the executed-reference fixtures.

Tolerances: boxes are single IEEE add/sub/clamp of identical fp32 inputs -> bit-exact.  Scores go through
sigmoid; the device's expf and ATen's (Sleef) differ by a few ulp, so scores are compared with rtol 2e-6
(written below) and the NMS stage is checked bit-exactly *at the NMS boundary*: the oracle is fed the very
candidates the GPU produced and must return identical indices and counts."""

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import canon
from test_oracle_fcos_match import CASES, load_fcos

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCORE_RTOL = 2e-6


def make_post(p, **kw):
    import types

    import oneshotdet_b200 as osd

    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(RPN_ONLY=False),
                                FEW_SHOT=types.SimpleNamespace(ADD_ARTIFICIAL_PROPOSALS=False))
    return osd.FCOSPostProcessor(cfg, p.pre_nms_thresh, p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n,
                                 p.min_size, num_classes=2, dense_points=1, score_calculator="BINARY", **kw).eval()


def to_dev(ts):
    return [t.to(DEV) for t in ts]


def gpu_candidates(res):
    """Compact per-episode candidate arrays from the slotted workspace layout."""
    cb, cs, cl, cnt, slot = res.candidates()
    cb, cs, cl, cnt = cb.cpu().numpy(), cs.cpu().numpy(), cl.cpu().numpy(), cnt.cpu().numpy()
    out = []
    for e in range(cb.shape[0]):
        bs, ss, ls, lv = [], [], [], []
        for l in range(cnt.shape[1]):
            n = cnt[e, l]
            bs.append(cb[e, slot[l]:slot[l] + n]); ss.append(cs[e, slot[l]:slot[l] + n])
            ls.append(cl[e, slot[l]:slot[l] + n]); lv.append(np.full(n, l))
        out.append((np.concatenate(bs), np.concatenate(ss), np.concatenate(lv), np.concatenate(ls)))
    return out


def check_against_oracle(cls, reg, ctr, sizes, p, early_exit=True, strict=False):
    post = make_post(p, early_exit=early_exit, strict_iou=strict)
    res = post.forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
    torch.cuda.synchronize()
    cands = gpu_candidates(res)
    oc = orc.fcos_candidates(cls, reg, ctr, orc.FPN_STRIDES, sizes, p)
    counts = res.count.cpu().numpy()
    for e, ((gb, gs, glv, gloc), (ob, os_, olv, oloc)) in enumerate(zip(cands, oc)):
        # --- candidate stage: same locations selected (score ulps could only matter exactly at a top-k boundary)
        same = gb.shape == ob.shape and np.array_equal(glv, olv) and np.array_equal(gloc, oloc)
        if same:
            np.testing.assert_array_equal(gb, ob)                       # decode + clip: bit-exact
            np.testing.assert_allclose(gs, os_, rtol=SCORE_RTOL, atol=0)  # sigmoid*sigmoid: few ulp
        else:
            # tolerate a swap at the top-k boundary only: the symmetric difference must be tiny and made of
            # scores within tolerance of the k-th score
            gset = set(zip(glv.tolist(), gloc.tolist())); oset = set(zip(olv.tolist(), oloc.tolist()))
            assert len(gset ^ oset) <= 4, f"episode {e}: candidate sets differ by {len(gset ^ oset)}"
        # --- NMS boundary: oracle on the GPU's own candidates -> identical indices, counts, order
        eb, es, ek = orc.select_over_all_levels(gb, gs, p, strict=strict)
        n = counts[e]
        assert n == ek.shape[0], f"episode {e}: kept {n} vs oracle {ek.shape[0]}"
        np.testing.assert_array_equal(res.index[e, :n].cpu().numpy(), ek)
        np.testing.assert_array_equal(res.boxes[e, :n].cpu().numpy(), eb)
        np.testing.assert_array_equal(res.scores[e, :n].cpu().numpy(), es)
    return res


@pytest.mark.parametrize("name", CASES)
def test_reference_fixtures(golden_dir, name):
    """Executed-reference outputs (tests/golden/make_golden.py) through the drop-in FCOSPostProcessor.forward."""
    z, cls, reg, ctr, p, sizes, outs = load_fcos(golden_dir, name)
    post = make_post(p)
    locs = [torch.from_numpy(z[f"loc{l}"]).to(DEV) for l in range(len(cls))]
    boxlists = post(locs, to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
    assert len(boxlists) == len(outs)
    for i, (bl, (rb, rs)) in enumerate(zip(boxlists, outs)):
        assert bl.mode == "xyxy" and bl.fields() == ["scores"]
        assert tuple(bl.size) == (sizes[i][1], sizes[i][0])
        gb, gs = bl.bbox.cpu().numpy(), bl.get_field("scores").cpu().numpy()
        assert gb.shape == rb.shape, (name, i, gb.shape, rb.shape)
        if rb.shape[0] == p.fpn_post_nms_top_n:
            np.testing.assert_array_equal(gb, rb)            # score-descending branch: order is part of the contract
            np.testing.assert_allclose(gs, rs, rtol=SCORE_RTOL, atol=0)
        else:
            gb, gs = canon(gb, gs); eb, es = canon(rb, rs)
            np.testing.assert_array_equal(gb, eb)
            np.testing.assert_allclose(gs, es, rtol=SCORE_RTOL, atol=0)


@pytest.mark.parametrize("name", CASES)
def test_fixture_inputs_vs_oracle_stages(golden_dir, name):
    z, cls, reg, ctr, p, sizes, outs = load_fcos(golden_dir, name)
    check_against_oracle(cls, reg, ctr, sizes, p)


def test_wrong_locations_are_rejected(golden_dir):
    z, cls, reg, ctr, p, sizes, outs = load_fcos(golden_dir, "nonms_small")
    post = make_post(p)
    locs = [torch.from_numpy(z[f"loc{l}"]).to(DEV) + 1.0 for l in range(len(cls))]
    with pytest.raises(ValueError, match="FCOS grid"):
        post(locs, to_dev(cls), to_dev(reg), to_dev(ctr), sizes)


@pytest.mark.parametrize("early_exit", [True, False])
def test_two_stage_800x1344(early_exit):
    """BASELINE config 1/2 geometry: 800x1333 padded to 800x1344, 11 600 candidates per episode."""
    cls, reg, ctr = orc.synth_head_outputs(2, 800, 1344, seed=2000)
    sizes = [(800, 1333), (800, 1333)]
    res = check_against_oracle(cls, reg, ctr, sizes, orc.TWO_STAGE, early_exit=early_exit)
    cnt = res.candidates()[3].cpu().numpy()
    np.testing.assert_array_equal(cnt, [[6000, 4200, 1050, 273, 77]] * 2)
    assert res.count.cpu().tolist() == [2000, 2000]
    kept = res.kept_before_cut().cpu().numpy()
    assert np.all(kept > 2000)


def test_early_exit_does_not_change_results():
    cls, reg, ctr = orc.synth_head_outputs(3, 512, 640, seed=2001)
    sizes = [(512, 640), (500, 600), (480, 640)]
    p = orc.PostParams(0.0, 1500, 0.7, 300, 0.0)
    a = make_post(p, early_exit=True).forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
    b = make_post(p, early_exit=False).forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
    assert torch.equal(a.count, b.count)
    for e, n in enumerate(a.count.tolist()):
        assert torch.equal(a.boxes[e, :n], b.boxes[e, :n]) and torch.equal(a.scores[e, :n], b.scores[e, :n])
        assert torch.equal(a.index[e, :n], b.index[e, :n])


def test_early_exit_miss_falls_through_to_second_pass():
    """Heavy suppression: fewer than post_top_n + 1 survivors among the first pass' rows, so pass 2 must run."""
    rng = np.random.RandomState(5)
    cls, reg, ctr = orc.synth_head_outputs(2, 512, 512, seed=2002)
    reg = [r * 0 + float(4 * s) for r, s in zip(reg, orc.FPN_STRIDES)]  # identical box shapes -> long chains
    sizes = [(512, 512)] * 2
    p = orc.PostParams(0.0, 2000, 0.3, 400, 0.0)
    check_against_oracle(cls, reg, ctr, sizes, p, early_exit=True)


@pytest.mark.parametrize("width", [12.0, 7.0])
def test_early_exit_passes_at_full_geometry_with_clustered_boxes(width):
    """BASELINE geometry with one box shape per level (2 x width strides wide): neighbouring locations suppress each other
    at IoU 0.8, the first pass (2112 rows) cannot reach post_top_n + 1 survivors and the later passes run -- with the
    kept-row compaction of their mask tiles.  The result must equal the single full pass (early_exit=False) bit for
    bit, and the oracle's keep list on the same candidates.  Stale workspace contents must not matter: the passes are
    run twice on different inputs through one FCOSPostProcessor (one workspace)."""
    sizes = [(800, 1333)] * 2
    post_e = make_post(orc.TWO_STAGE, early_exit=True)
    post_f = make_post(orc.TWO_STAGE, early_exit=False)
    for seed in (2100, 2101):
        cls, reg, ctr = orc.synth_head_outputs(2, 800, 1344, seed=seed)
        reg = [r * 0 + float(width * s) for r, s in zip(reg, orc.FPN_STRIDES)]
        a = post_e.forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
        b = post_f.forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), sizes)
        torch.cuda.synchronize()
        assert torch.equal(a.count, b.count)
        for e, n in enumerate(a.count.tolist()):
            assert torch.equal(a.index[e, :n], b.index[e, :n]) and torch.equal(a.boxes[e, :n], b.boxes[e, :n])
            assert torch.equal(a.scores[e, :n], b.scores[e, :n])
        # the exit must really have been missed in pass 1 for this test to mean anything
        kept = a.kept_before_cut().cpu().numpy()
        full = b.kept_before_cut().cpu().numpy()
        assert np.all(full < 0.9 * 11600), full
    check_against_oracle(cls, reg, ctr, sizes, orc.TWO_STAGE, early_exit=True)


def test_stress_params_config5():
    """BASELINE config 5: thresh 0.01, 1000 per level, NMS 0.6, 20 episodes sharing one image size."""
    cls, reg, ctr = orc.synth_head_outputs(4, 800, 1344, seed=5000)
    check_against_oracle(cls, reg, ctr, [(800, 1333)] * 4, orc.STRESS)


def test_config4_1024_square():
    cls, reg, ctr = orc.synth_head_outputs(2, 1024, 1024, seed=4000)
    check_against_oracle(cls, reg, ctr, [(1024, 1024)] * 2, orc.TWO_STAGE)


def test_strict_iou_variant():
    cls, reg, ctr = orc.synth_head_outputs(2, 256, 384, seed=2003)
    check_against_oracle(cls, reg, ctr, [(256, 384)] * 2, orc.PostParams(0.02, 500, 0.5, 100, 0.0), strict=True)


def test_min_size_and_no_candidates():
    cls, reg, ctr = orc.synth_head_outputs(2, 256, 256, seed=2004)
    p = orc.PostParams(0.05, 400, 0.6, 100, 40.0)
    check_against_oracle(cls, reg, ctr, [(200, 256), (256, 256)], p)
    # nothing above the threshold in episode 1
    cls = [c.clone() for c in cls]
    for c in cls:
        c[1] = -50.0
    res = check_against_oracle(cls, reg, ctr, [(200, 256), (256, 256)], p)
    assert res.count.cpu().tolist()[1] == 0


def test_tied_scores_at_topk_boundary_are_deterministic():
    """Quantised logits: many equal scores; selection must be 'lowest location first', NMS order stable."""
    cls, reg, ctr = orc.synth_head_outputs(2, 256, 384, seed=2005, distinct=False)
    cls = [torch.round(c * 4) / 4 for c in cls]
    ctr = [torch.zeros_like(c) for c in ctr]
    check_against_oracle(cls, reg, ctr, [(256, 384)] * 2, orc.PostParams(0.0, 200, 0.6, 150, 0.0))


def test_multi_class_logits_use_channel_zero():
    """num_classes > 2 with BINARY scores keeps the positive channel only (inference.py:57-61)."""
    import types

    import oneshotdet_b200 as osd

    cls, reg, ctr = orc.synth_head_outputs(1, 128, 128, seed=2006)
    p = orc.PostParams(0.02, 100, 0.5, 50, 0.0)
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(RPN_ONLY=True))
    post = osd.FCOSPostProcessor(cfg, p.pre_nms_thresh, p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n, p.min_size,
                                 num_classes=3, dense_points=1, score_calculator="BINARY").eval()
    two = [torch.cat((c, torch.randn_like(c)), dim=1) for c in cls]
    a = post.forward_fixed(to_dev(two), to_dev(reg), to_dev(ctr), [(128, 128)])
    b = make_post(p).forward_fixed(to_dev(cls), to_dev(reg), to_dev(ctr), [(128, 128)])
    n = int(a.count[0])
    assert torch.equal(a.count, b.count) and torch.equal(a.boxes[:, :n], b.boxes[:, :n])
    assert torch.equal(a.scores[:, :n], b.scores[:, :n])


def test_raw_regression_with_folded_exp_scale():
    """box_regression given as the RAW bbox_pred conv output plus the per-level Scale parameters;
    the head's tail exp(x * scale_l) (fcos.py:95-97) is evaluated inside the decode.  Tolerance: the device expf and
    ATen's CPU exp differ by <= 2 ulp of the distance (<= 2.4e-4 px at distances < 2048 px) and the box coordinate
    adds one rounding -> |delta| <= 1e-3 px, written below; selection, scores and the NMS boundary stay exact."""
    b, h, w = 3, 200, 264
    cls, reg, ctr = orc.synth_head_outputs(b, h, w, seed=31)
    scales = [1.0, 0.83, 1.21, 0.97, 1.4]
    raw = [torch.log(r) / s for r, s in zip(reg, scales)]
    sizes = [(200, 260), (190, 264), (200, 264)]
    p = orc.PostParams(0.0, 800, 0.7, 150, 0.0)
    post = make_post(p)
    res = post.forward_fixed(to_dev(cls), to_dev(raw), to_dev(ctr), sizes, reg_scales=scales)
    torch.cuda.synchronize()
    reg_ref = [orc.fcos_head_tail(x, s) for x, s in zip(raw, scales)]     # the reference's elementwise passes
    oc = orc.fcos_candidates(cls, reg_ref, ctr, orc.FPN_STRIDES, sizes, p)
    counts = res.count.cpu().numpy()
    cands = gpu_candidates(res)                                            # (the workspace is shared between calls)
    for e, ((gb, gs, glv, gloc), (ob, os_, olv, oloc)) in enumerate(zip(cands, oc)):
        np.testing.assert_array_equal(glv, olv)
        np.testing.assert_array_equal(gloc, oloc)
        np.testing.assert_allclose(gs, os_, rtol=SCORE_RTOL, atol=0)
        np.testing.assert_allclose(gb, ob, rtol=0, atol=1e-3)
        eb, es, ek = orc.select_over_all_levels(gb, gs, p)                 # NMS boundary on the GPU's candidates
        n = counts[e]
        assert n == ek.shape[0]
        np.testing.assert_array_equal(res.index[e, :n].cpu().numpy(), ek)
        np.testing.assert_array_equal(res.boxes[e, :n].cpu().numpy(), eb)
    # and with scale 1 on already-exp'd input the two entry forms agree to the same tolerance
    res2 = post.forward_fixed(to_dev(cls), to_dev(reg_ref), to_dev(ctr), sizes)
    torch.cuda.synchronize()
    for (gb, *_), (hb, *_) in zip(cands, gpu_candidates(res2)):
        np.testing.assert_allclose(gb, hb, rtol=0, atol=1e-3)
