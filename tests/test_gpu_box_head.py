"""GPU: the built halves of the second stage (Pooler kernel, PostProcessor kernel) against the fixture produced by
executing the reference's whole ROIBoxHead.forward; the dense middle (row 2c) is evaluated by the oracle here until its
kernels exist."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import canon
from test_oracle_box_head import load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pooler_and_postprocessor_reproduce_the_reference_second_stage(golden_dir):
    import types

    import oneshotdet_b200 as osd

    z, mods = load(golden_dir)
    b, c, h, w = int(z["batch"]), int(z["channels"]), int(z["height"]), int(z["width"])
    feats, _ = orc.synth_features(b, 1, c, h, w, int(z["seed"]))
    sizes = [tuple(int(v) for v in hw) for hw in z["image_sizes"]]
    boxes = torch.from_numpy(z["boxes"])
    # 1. pooling kernel == the reference's FPN2ROIFeatureExtractor output, bit for bit
    pooler = osd.Pooler((7, 7), [1 / s for s in orc.FPN_STRIDES], 2)
    bl = [osd.BoxList(boxes[i].to(DEV), (sizes[i][1], sizes[i][0]), mode="xyxy") for i in range(b)]
    pooled = pooler([f.to(DEV) for f in feats], bl)
    np.testing.assert_array_equal(pooled.cpu().numpy(), z["pooled"])
    # 2. dense middle: oracle (no kernel yet) on the kernel's pooled features
    logits, reg = orc.box_head_dense(pooled.cpu(), torch.from_numpy(z["supp"]), mods)
    np.testing.assert_allclose(logits.numpy(), z["class_logits"], rtol=1e-5, atol=1e-6)
    # 3. post-processing kernel on those logits == the reference's final BoxLists
    st, nt, dpi = z["params"]
    cfg = types.SimpleNamespace(FEW_SHOT=types.SimpleNamespace(SECOND_STAGE_CLS_LOSS="ce_loss"))
    post = osd.PostProcessor(cfg, float(st), float(nt), int(dpi), osd.BoxCoder(tuple(float(v) for v in z["weights"])), False).eval()
    out = post((logits.to(DEV), reg.to(DEV)), bl, target_ids=z["target_ids"].tolist())
    for i, o in enumerate(out):
        rb, rs = z[f"out_boxes{i}"], z[f"out_scores{i}"]
        gb, gs = o.bbox.cpu().numpy(), o.get_field("scores").cpu().numpy()
        assert gb.shape == rb.shape
        gb, gs = canon(gb, gs); rb, rs = canon(rb, rs)
        np.testing.assert_allclose(gs, rs, rtol=2e-5, atol=0)     # expf vs ATen exp on logits that differ by a few ulp
        np.testing.assert_allclose(gb, rb, rtol=0, atol=2e-3)
        np.testing.assert_array_equal(o.get_field("labels").cpu().numpy(), z[f"out_labels{i}"][:len(gs)])


# ------------------------------------------------------------------------------------------------------
# the dense middle on the GPU (csrc/box_head.cu): six tcgen05 GEMM launches with GroupNorm / LeakyReLU epilogues
# ------------------------------------------------------------------------------------------------------
def _bf(t):
    return t.bfloat16().float()


def emulated_dense(pooled, supp, m):
    """The oracle's box_head_dense on bf16-ROUNDED operands and bf16 activations between layers, fp32 accumulation --
    what the tensor-core path computes (differences left: summation order, one-pass GroupNorm variance)."""
    import torch.nn.functional as F

    b, r, c, p, _ = pooled.shape
    cdc, agg = m["compress_dim_conv"], m["feature_aggreg"]
    with torch.no_grad():
        x = torch.cat((_bf(pooled).reshape(-1, c, p, p), _bf(supp).expand_as(pooled).reshape(-1, c, p, p)), 1)
        y = F.conv2d(x, _bf(cdc[0].weight), cdc[0].bias)
        a1 = _bf(cdc[2](cdc[1](y)))
        y = F.conv2d(a1, _bf(cdc[3].weight), cdc[3].bias)
        a2 = _bf(cdc[5](cdc[4](y)))
        y = F.conv2d(a2, _bf(agg[0].weight), agg[0].bias, padding=1)
        a3 = _bf(agg[2](agg[1](y)))
        h6 = _bf(F.relu(F.linear(a3.reshape(a3.size(0), -1), _bf(m["fc6"].weight), m["fc6"].bias)))
        h7 = _bf(F.relu(F.linear(h6, _bf(m["fc7"].weight), m["fc7"].bias)))
        return (F.linear(h7, _bf(m["cls_score"].weight), m["cls_score"].bias),
                F.linear(h7, _bf(m["bbox_pred"].weight), m["bbox_pred"].bias), (a1, a2, a3, h6, h7))


def gpu_head(mods, c, mlp, roi_chunk=0):
    import oneshotdet_b200 as osd

    head = osd.BoxHeadDense(c, mlp, roi_chunk=roi_chunk)
    sd = {("predictor." + k if k.startswith(("cls_score", "bbox_pred")) else k): v for k, v in mods.state_dict().items()}
    head.load_state_dict(sd)
    return head.to(DEV).eval()


def rel_err(got, ref):
    return float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)


def test_dense_head_kernels_against_the_executed_reference(golden_dir):
    """class_logits / box_regression of the reference's executed ROIBoxHead.forward (fp32) from the tcgen05 path
    (bf16 operands and activations, fp32 accumulation): within 2 % of the output range; within 0.5 % of the oracle
    evaluated on bf16-rounded operands.  Then the whole second stage on the GPU: pooler -> dense head -> post-processor."""
    import types

    import oneshotdet_b200 as osd

    z, mods = load(golden_dir)
    c, mlp = int(z["channels"]), int(z["w_fc7.weight"].shape[0])
    pooled, supp = torch.from_numpy(z["pooled"]), torch.from_numpy(z["supp"])
    head = gpu_head(mods, c, mlp)
    logits, reg = head(pooled.to(DEV), supp.to(DEV))
    torch.cuda.synchronize()
    logits, reg = logits.cpu(), reg.cpu()
    el, er, _ = emulated_dense(pooled, supp, mods)
    assert rel_err(logits, el) <= 5e-3 and rel_err(reg, er) <= 5e-3, (rel_err(logits, el), rel_err(reg, er))
    rl, rr = torch.from_numpy(z["class_logits"]), torch.from_numpy(z["box_regression"])
    assert rel_err(logits, rl) <= 2e-2 and rel_err(reg, rr) <= 2e-2, (rel_err(logits, rl), rel_err(reg, rr))

    # the whole second stage on the GPU
    b, h, w = int(z["batch"]), int(z["height"]), int(z["width"])
    feats, _ = orc.synth_features(b, 1, c, h, w, int(z["seed"]))
    sizes = [tuple(int(v) for v in hw) for hw in z["image_sizes"]]
    boxes = torch.from_numpy(z["boxes"])
    pooler = osd.Pooler((7, 7), [1 / s for s in orc.FPN_STRIDES], 2)
    bl = [osd.BoxList(boxes[i].to(DEV), (sizes[i][1], sizes[i][0]), mode="xyxy") for i in range(b)]
    lg, rg = head(pooler([f.to(DEV) for f in feats], bl), supp.to(DEV))
    st, nt, dpi = z["params"]
    cfg = types.SimpleNamespace(FEW_SHOT=types.SimpleNamespace(SECOND_STAGE_CLS_LOSS="ce_loss"))
    post = osd.PostProcessor(cfg, float(st), float(nt), int(dpi), osd.BoxCoder(tuple(float(v) for v in z["weights"])), False).eval()
    out = post((lg, rg), bl, target_ids=z["target_ids"].tolist())
    for i, o in enumerate(out):
        rb, rs = z[f"out_boxes{i}"], z[f"out_scores{i}"]
        gb, gs = o.bbox.cpu().numpy(), o.get_field("scores").cpu().numpy()
        # scores are softmax probabilities of logits that differ by bf16 rounding: the detections above the score
        # threshold are the same set unless a score sits within that error of the threshold
        assert abs(gb.shape[0] - rb.shape[0]) <= 1
        if gb.shape == rb.shape:
            gb, gs = canon(gb, gs); rb, rs = canon(rb, rs)
            np.testing.assert_allclose(gs, rs, rtol=0, atol=5e-3)


@pytest.mark.parametrize("c,mlp,b,r,chunk", [(64, 64, 2, 5, 0), (128, 256, 1, 9, 4), (256, 1024, 2, 37, 16), (256, 1024, 3, 64, 0)])
def test_dense_head_layers_against_bf16_emulation(c, mlp, b, r, chunk):
    """Random weights and inputs at the reference's widths, odd ROI counts and several chunks: every layer's output
    (read back from the workspace when one chunk holds all ROIs) and the final outputs against the bf16-emulated oracle."""
    torch.manual_seed(100 + c + r)
    mods = orc.make_box_head_modules(c, mlp)
    with torch.no_grad():
        for p_ in mods.parameters():            # default inits give near-zero logits: use O(1)-signal weights
            if p_.dim() > 1:
                p_.normal_(std=1.0 / (p_[0].numel() ** 0.5))
            else:
                p_.uniform_(-0.5, 0.5)
        for k in ("compress_dim_conv.1", "compress_dim_conv.4", "feature_aggreg.1"):
            mods.get_submodule(k).weight.uniform_(0.5, 1.5)
    pooled = torch.randn(b, r, c, 7, 7)
    supp = torch.randn(b, 1, c, 7, 7)
    head = gpu_head(mods, c, mlp, roi_chunk=chunk)
    logits, reg = head(pooled.to(DEV), supp.to(DEV))
    torch.cuda.synchronize()
    el, er, _ = emulated_dense(pooled, supp, mods)
    # a bf16 rounding boundary crossed in one activation moves later layers by up to one bf16 ulp of that value
    assert rel_err(logits.cpu(), el) <= 1e-2, rel_err(logits.cpu(), el)
    assert rel_err(reg.cpu(), er) <= 1e-2, rel_err(reg.cpu(), er)
    fl, fr = orc.box_head_dense(pooled, supp, mods)
    assert rel_err(logits.cpu(), fl) <= 3e-2 and rel_err(reg.cpu(), fr) <= 3e-2


def test_pooler_bf16_rows_feed_the_dense_head(golden_dir):
    """Pooler(rows_bf16=True) writes the pooled features once, as bf16 [B,R,49,C] rows: the fp32 result rounded to bf16
    and transposed, bit for bit; the dense head on them equals the dense head on the fp32 tensor (same rounding, same
    GEMMs) exactly."""
    import oneshotdet_b200 as osd

    z, mods = load(golden_dir)
    b, c, h, w = int(z["batch"]), int(z["channels"]), int(z["height"]), int(z["width"])
    feats, _ = orc.synth_features(b, 1, c, h, w, int(z["seed"]))
    rois = torch.from_numpy(z["boxes"]).to(DEV)
    pooler = osd.Pooler((7, 7), [1 / s for s in orc.FPN_STRIDES], 2)
    fx = [f.to(DEV) for f in feats]
    rows = pooler.forward_fixed(fx, rois, rows_bf16=True)
    full = pooler.forward_fixed(fx, rois)
    assert rows.dtype == torch.bfloat16 and tuple(rows.shape) == (b, rois.size(1), 49, c)
    want = full.reshape(b, rois.size(1), c, 49).permute(0, 1, 3, 2).bfloat16()
    assert torch.equal(rows, want)
    head = gpu_head(mods, c, int(z["w_fc7.weight"].shape[0]))
    supp = torch.from_numpy(z["supp"]).to(DEV)
    l0, r0 = head(full, supp)
    l1, r1 = head(rows, supp)
    assert torch.equal(l0, l1) and torch.equal(r0, r1)
    # padded ROI lists (roi_count): rows past an image's count are zeros in both layouts
    cnt = torch.tensor([rois.size(1) - 5, rois.size(1)], dtype=torch.int32, device=DEV)[:b]
    rows_c = pooler.forward_fixed(fx, rois, cnt, rows_bf16=True)
    full_c = pooler.forward_fixed(fx, rois, cnt)
    assert torch.equal(rows_c, full_c.reshape(b, rois.size(1), c, 49).permute(0, 1, 3, 2).bfloat16())
    assert float(rows_c[0, rois.size(1) - 5:].float().abs().max()) == 0.0 and torch.equal(rows_c[1], rows[1])


def test_dense_head_is_independent_of_chunking_and_roi_order():
    """Size-independent properties at a size the CPU oracle does not reach (4 x 700 ROIs, C = 256, MLP 1024): every ROI's
    result depends on that ROI and its episode's support only -- so the outputs are bit-identical whatever the chunk
    size (which ROIs share a launch / a 128-row tile) and a permutation of the ROIs inside an episode permutes the
    outputs."""
    torch.manual_seed(11)
    c, mlp, b, r = 256, 1024, 4, 700
    mods = orc.make_box_head_modules(c, mlp)
    pooled = torch.randn(b, r, c, 7, 7, device=DEV)
    supp = torch.randn(b, 1, c, 7, 7, device=DEV)
    ref_l, ref_r = gpu_head(mods, c, mlp)(pooled, supp)
    for chunk in (2, 64, 1001):
        l, rg = gpu_head(mods, c, mlp, roi_chunk=chunk)(pooled, supp)
        assert torch.equal(l, ref_l) and torch.equal(rg, ref_r), chunk
    perm = torch.randperm(r, device=DEV)
    lp, rp = gpu_head(mods, c, mlp)(pooled[:, perm].contiguous(), supp)
    assert torch.equal(lp.view(b, r, -1), ref_l.view(b, r, -1)[:, perm])
    assert torch.equal(rp.view(b, r, -1), ref_r.view(b, r, -1)[:, perm])
    assert bool(torch.isfinite(ref_l).all()) and float(ref_l.abs().max()) > 0
