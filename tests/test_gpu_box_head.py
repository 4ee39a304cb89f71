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
