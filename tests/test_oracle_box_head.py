"""The whole second stage restated by the oracle -- Pooler -> concat + compress_dim_conv -> feature_aggreg -> fc6/fc7 ->
FPNPredictor -> PostProcessor -- against a fixture produced by EXECUTING the reference's ROIBoxHead.forward (with its own
config defaults) in the build container.  The dense middle has no kernel yet; this pins the oracle
the kernels will be held to."""
import os

import numpy as np
import torch

from oracle import oracle as orc
from helpers import canon


def load(golden_dir, name="c64"):
    z = np.load(os.path.join(golden_dir, f"box_head_{name}.npz"))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    c, mlp = int(z["channels"]), int(z["w_fc7.weight"].shape[0])
    return z, orc.make_box_head_modules(c, mlp, state=state)


def test_pooled_features_match_the_reference_pooler(golden_dir):
    z, _ = load(golden_dir)
    b, c, h, w = int(z["batch"]), int(z["channels"]), int(z["height"]), int(z["width"])
    feats, _ = orc.synth_features(b, 1, c, h, w, int(z["seed"]))
    out, _ = orc.pooler_forward(feats, torch.from_numpy(z["boxes"]), [1 / s for s in orc.FPN_STRIDES], 7, 2)
    np.testing.assert_array_equal(out.reshape(z["pooled"].shape), z["pooled"])


def test_dense_head_stages_match_the_executed_reference(golden_dir):
    z, mods = load(golden_dir)
    pooled, supp = torch.from_numpy(z["pooled"]), torch.from_numpy(z["supp"])
    # same ATen kernels on the same inputs: equal up to oneDNN's choice of blocking (allow a few ulp)
    for stage, key in (("compressed", "compressed"), ("aggregated", "aggregated"), ("fc7", "fc7_pre_relu")):
        got = orc.box_head_dense(pooled, supp, mods, stage).numpy()
        np.testing.assert_allclose(got, z[key], rtol=1e-5, atol=1e-6)
    logits, reg = orc.box_head_dense(pooled, supp, mods)
    np.testing.assert_allclose(logits.numpy(), z["class_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(reg.numpy(), z["box_regression"], rtol=1e-5, atol=1e-6)


def test_whole_second_stage_detections(golden_dir):
    z, mods = load(golden_dir)
    st, nt, dpi = z["params"]
    p = orc.BoxPostParams(float(st), float(nt), int(dpi), tuple(float(v) for v in z["weights"]), "softmax")
    sizes = [tuple(int(v) for v in hw) for hw in z["image_sizes"]]
    # post-processing on the reference's own logits: exact
    res = orc.box_postprocess(torch.from_numpy(z["class_logits"]), torch.from_numpy(z["box_regression"]),
                              torch.from_numpy(z["boxes"]), sizes, p)
    for i, r in enumerate(res):
        assert r["boxes"].shape == z[f"out_boxes{i}"].shape
        gb, gs = canon(r["boxes"], r["scores"]); eb, es = canon(z[f"out_boxes{i}"], z[f"out_scores{i}"])
        np.testing.assert_array_equal(gb, eb)
        np.testing.assert_array_equal(gs, es)
        assert (z[f"out_labels{i}"] == z["target_ids"][i]).all()
