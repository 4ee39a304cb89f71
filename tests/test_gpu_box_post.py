"""GPU parity of the second-stage post-processing (osd_box_postprocess through the reference-shaped PostProcessor)
against the oracle and the executed-reference fixtures.

Tolerances: the class probability and the decoded width/height go through expf (device) vs ATen's CPU exp: a few ulp.
Scores are compared with rtol 2e-6; box coordinates with atol 2e-3 px (exp(dw) * w up to ~4000 px: 2 ulp = 5e-4, plus
the roundings of the following adds).  The NMS stage is checked bit-exactly at the NMS boundary: the oracle is fed the
candidates the GPU produced and must return identical indices, counts and order."""
import types

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import canon
from test_oracle_box_post import BOX_CASES, load_box_post

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCORE_RTOL = 2e-6
BOX_ATOL = 2e-3


def make_post(p, cls_loss="ce_loss", agnostic=False):
    import oneshotdet_b200 as osd

    cfg = types.SimpleNamespace(FEW_SHOT=types.SimpleNamespace(SECOND_STAGE_CLS_LOSS=cls_loss))
    return osd.PostProcessor(cfg, p.score_thresh, p.nms_thresh, p.detections_per_img, osd.BoxCoder(p.weights), agnostic).eval()


def gpu_cands(res):
    cb, cs, src, cnt = (t.cpu().numpy() for t in res.candidates())
    return [(cb[i, :cnt[i]], cs[i, :cnt[i]], src[i, :cnt[i]]) for i in range(cb.shape[0])]


def check(logits, reg, props, sizes, p, cls_loss="ce_loss", roi_count=None):
    post = make_post(p, cls_loss)
    rc = None if roi_count is None else torch.tensor(roi_count, dtype=torch.int32, device=DEV)
    res = post.forward_fixed((logits.to(DEV), reg.to(DEV)), props.to(DEV), sizes, roi_count=rc)
    torch.cuda.synchronize()
    counts = res.count.cpu().numpy()
    oc = orc.box_candidates(logits, reg, props, sizes, p, roi_count)
    for i, ((gb, gs, gsrc), (ob, os_, osrc)) in enumerate(zip(gpu_cands(res), oc)):
        np.testing.assert_array_equal(gsrc, osrc)                        # same proposals pass the threshold, same order
        np.testing.assert_allclose(gs, os_, rtol=SCORE_RTOL, atol=0)
        np.testing.assert_allclose(gb, ob, rtol=0, atol=BOX_ATOL)
        eb, es, ek = orc.box_filter_results(gb, gs, p)                   # NMS boundary on the GPU's own candidates
        n = counts[i]
        assert n == ek.shape[0], f"image {i}: kept {n} vs oracle {ek.shape[0]}"
        np.testing.assert_array_equal(res.index[i, :n].cpu().numpy(), ek)
        np.testing.assert_array_equal(res.boxes[i, :n].cpu().numpy(), eb)
        np.testing.assert_array_equal(res.scores[i, :n].cpu().numpy(), es)
    return res


@pytest.mark.parametrize("name", BOX_CASES)
def test_reference_fixtures(golden_dir, name):
    """Executed-reference outputs through the drop-in PostProcessor.forward (BoxLists in, BoxLists out)."""
    import oneshotdet_b200 as osd

    z, logits, reg, props, p, sizes, outs = load_box_post(golden_dir, name)
    post = make_post(p, str(z["cls_loss"]), bool(int(z["agnostic"])))
    boxes = [osd.BoxList(props[i].to(DEV), (sizes[i][1], sizes[i][0]), mode="xyxy") for i in range(len(sizes))]
    out = post((logits.to(DEV), reg.to(DEV)), boxes, target_ids=z["target_ids"].tolist())
    assert len(out) == len(outs)
    for i, (bl, (rb, rs, rl)) in enumerate(zip(out, outs)):
        assert bl.mode == "xyxy" and sorted(bl.fields()) == ["labels", "scores"]
        assert tuple(bl.size) == (sizes[i][1], sizes[i][0])
        gb, gs = bl.bbox.cpu().numpy(), bl.get_field("scores").cpu().numpy()
        assert gb.shape == rb.shape, (name, i, gb.shape, rb.shape)
        np.testing.assert_array_equal(bl.get_field("labels").cpu().numpy(), rl)
        if rb.shape[0] != p.detections_per_img:
            gb, gs = canon(gb, gs); rb, rs = canon(rb, rs)
        np.testing.assert_allclose(gs, rs, rtol=SCORE_RTOL, atol=0)
        np.testing.assert_allclose(gb, rb, rtol=0, atol=BOX_ATOL)


@pytest.mark.parametrize("name", BOX_CASES)
def test_fixture_inputs_vs_oracle_stages(golden_dir, name):
    z, logits, reg, props, p, sizes, outs = load_box_post(golden_dir, name)
    check(logits, reg, props, sizes, p, str(z["cls_loss"]))


def test_full_size_two_stage_defaults():
    """The shipped configuration: 2000 proposals per image, SCORE_THRESH 0, NMS 0.5, 2000 detections, 16 images."""
    sizes = [(800, 1333)] * 16
    logits, reg, props = orc.synth_box_head_outputs(16, 2000, sizes, seed=51)
    res = check(logits, reg, props, sizes, orc.BoxPostParams())
    assert 0 < int(res.count.min()) and int(res.count.max()) < 2000


def test_ragged_roi_counts_and_many_logits():
    sizes = [(300, 500), (310, 480), (128, 128)]
    logits, reg, props = orc.synth_box_head_outputs(3, 500, sizes, seed=52, num_logits=5, reg_columns=12)
    check(logits, reg, props, sizes, orc.BoxPostParams(0.1, 0.6, 30), roi_count=[500, 123, 0])


def test_chained_from_fcos_stage_outputs():
    """Proposals straight from the FCOS stage's padded device output ([B,K,4] + counts): no host sync in between."""
    import oneshotdet_b200 as osd

    b, h, w = 2, 256, 320
    cls, regm, ctr = orc.synth_head_outputs(b, h, w, seed=61)
    sizes = [(250, 320), (256, 300)]
    fp = orc.PostParams(0.0, 500, 0.7, 120, 0.0)
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(RPN_ONLY=False))
    fpost = osd.FCOSPostProcessor(cfg, fp.pre_nms_thresh, fp.pre_nms_top_n, fp.nms_thresh, fp.fpn_post_nms_top_n,
                                  fp.min_size, 2, 1, "BINARY").eval()
    first = fpost.forward_fixed([t.to(DEV) for t in cls], [t.to(DEV) for t in regm], [t.to(DEV) for t in ctr], sizes)
    k = first.boxes.size(1)
    rng = np.random.RandomState(7)
    logits = torch.from_numpy(rng.normal(0, 2, (b * k, 2)).astype(np.float32))
    reg = torch.from_numpy((rng.normal(0, 0.3, (b * k, 8))).astype(np.float32))
    p = orc.BoxPostParams(0.0, 0.5, 2000)
    res = make_post(p).forward_fixed((logits.to(DEV), reg.to(DEV)), first.boxes, sizes, roi_count=first.count)
    torch.cuda.synchronize()
    cnt = first.count.cpu().tolist()
    oc = orc.box_candidates(logits, reg, first.boxes.cpu(), sizes, p, roi_count=cnt)
    for i, ((gb, gs, gsrc), (ob, os_, osrc)) in enumerate(zip(gpu_cands(res), oc)):
        np.testing.assert_array_equal(gsrc, osrc)
        np.testing.assert_allclose(gb, ob, rtol=0, atol=BOX_ATOL)
        eb, es, ek = orc.box_filter_results(gb, gs, p)
        n = int(res.count[i].item())
        assert n == ek.shape[0]
        np.testing.assert_array_equal(res.index[i, :n].cpu().numpy(), ek)


def test_bad_arguments_raise():
    import oneshotdet_b200 as osd

    p = orc.BoxPostParams()
    post = make_post(p)
    with pytest.raises(ValueError):
        post.forward_fixed((torch.zeros(4, 2, device=DEV), torch.zeros(4, 4, device=DEV)), torch.zeros(1, 4, 4, device=DEV), [(8, 8)])
    with pytest.raises(RuntimeError):
        osd.box_postprocess(torch.zeros(4, 2, device=DEV), torch.zeros(3, 8, device=DEV), torch.zeros(1, 4, 4, device=DEV), [(8, 8)])
    with pytest.raises(RuntimeError):
        osd.box_postprocess(torch.zeros(4, 2), torch.zeros(4, 8), torch.zeros(1, 4, 4), [(8, 8)])
