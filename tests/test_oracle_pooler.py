"""The oracle's restatement of the second-stage Pooler (LevelMapper + ROIAlign PxP) against (1) fixtures produced by
EXECUTING the reference's modeling/poolers.py with its compiled ROIAlign_cpu.cpp, (2) the reference's compiled operator
itself when oracle/_ref is present."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

POOLER_CASES = ["r7_s2", "all_levels", "r3_adaptive"]


def load_pooler(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"pooler_{name}.npz"))
    b, c, h, w = int(z["batch"]), int(z["channels"]), int(z["height"]), int(z["width"])
    feats, _ = orc.synth_features(b, 1, c, h, w, int(z["seed"]))
    boxes = torch.from_numpy(z["boxes"])
    return z, feats, boxes, [float(s) for s in z["scales"]], int(z["resolution"]), int(z["sampling"])


@pytest.mark.parametrize("name", POOLER_CASES)
def test_oracle_matches_executed_reference(golden_dir, name):
    z, feats, boxes, scales, res, samp = load_pooler(golden_dir, name)
    np.testing.assert_array_equal(boxes.numpy(), orc.synth_rois(int(z["batch"]), int(z["rois"]),
                                                                [tuple(v) for v in z["image_sizes"]],
                                                                int(z["seed"]) + 1, float(z["lo"])).numpy())
    out, levels = orc.pooler_forward(feats, boxes, scales, res, samp)
    np.testing.assert_array_equal(levels, z["levels"])
    want = z["pooled"].reshape(out.shape)
    np.testing.assert_array_equal(out, want)          # bit-exact: same operations in the same order


def test_restatement_equals_compiled_reference_operator(ref_ops):
    if ref_ops is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    sizes = [(300, 420), (280, 400)]
    feats, _ = orc.synth_features(2, 1, 6, 320, 448, seed=5)
    rois = orc.synth_rois(2, 24, sizes, 6)
    scales = [1 / s for s in orc.FPN_STRIDES]
    fn = lambda f, q, s, p, g: ref_ops.roi_align_forward(f.contiguous(), q.contiguous(), s, p, p, g)  # noqa: E731
    for res, samp in [(7, 2), (2, 3), (5, 0)]:
        a, la = orc.pooler_forward(feats, rois, scales, res, samp)
        b, lb = orc.pooler_forward(feats, rois, scales, res, samp, roi_align_fn=fn)
        np.testing.assert_array_equal(la, lb)
        np.testing.assert_array_equal(a, b)


def test_level_mapper_boundaries():
    """floor(4 + log2(sqrt(area)/224 + 1e-6)) clamped to [3, 7]: canonical 224 px boxes sit on level 4 (offset 1)."""
    def box(s):
        return [0.0, 0.0, s - 1.0, s - 1.0]        # legacy +1 widths: side s
    b = torch.tensor([box(10), box(111), box(112), box(223), box(224), box(447), box(448), box(1791), box(1792), box(5000)])
    np.testing.assert_array_equal(orc.map_levels(b, 3.0, 7.0).numpy(), [0, 0, 0, 0, 1, 1, 2, 3, 4, 4])
