"""The oracle's restatement of the second-stage PostProcessor (oracle.box_*) against fixtures produced by EXECUTING
the reference's modeling/roi_heads/box_head/inference.py + modeling/box_coder.py (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import canon

BOX_CASES = ["default", "cut_thresh", "focal_agnostic"]


def load_box_post(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"box_post_{name}.npz"))
    st, nt, dpi = z["params"]
    mode = "sigmoid" if str(z["cls_loss"]) == "focal_loss" else "softmax"
    p = orc.BoxPostParams(float(st), float(nt), int(dpi), tuple(float(w) for w in z["weights"]), mode)
    sizes = [tuple(int(v) for v in hw) for hw in z["image_sizes"]]
    outs = [(z[f"out_boxes{i}"], z[f"out_scores{i}"], z[f"out_labels{i}"]) for i in range(int(z["batch"]))]
    return z, torch.from_numpy(z["logits"]), torch.from_numpy(z["reg"]), torch.from_numpy(z["props"]), p, sizes, outs


@pytest.mark.parametrize("name", BOX_CASES)
def test_oracle_matches_executed_reference(golden_dir, name):
    z, logits, reg, props, p, sizes, outs = load_box_post(golden_dir, name)
    res = orc.box_postprocess(logits, reg, props, sizes, p)
    assert len(res) == len(outs)
    for i, (r, (rb, rs, rl)) in enumerate(zip(res, outs)):
        assert r["boxes"].shape == rb.shape, (name, i)
        assert sorted(str(f) for f in z[f"out_fields{i}"]) == ["labels", "scores"]
        assert np.all(rl == z["target_ids"][i])
        assert tuple(z[f"out_size{i}"]) == (sizes[i][1], sizes[i][0])
        if rb.shape[0] == p.detections_per_img:        # score-descending branch: the order is part of the contract
            np.testing.assert_array_equal(r["boxes"], rb)
            np.testing.assert_array_equal(r["scores"], rs)
        else:
            gb, gs = canon(r["boxes"], r["scores"]); eb, es = canon(rb, rs)
            np.testing.assert_array_equal(gb, eb)
            np.testing.assert_array_equal(gs, es)


def test_decode_clamps_large_deltas_and_keeps_legacy_plus_one():
    boxes = torch.tensor([[10., 20., 29., 59.]])                 # w = 20, h = 40 (legacy +1)
    codes = torch.tensor([[0., 0., 0., 0., 1.0, -2.0, 100.0, 0.0]])
    out = orc.box_decode(codes, boxes, (10., 10., 5., 5.))
    np.testing.assert_allclose(out[0, :4].numpy(), [10., 20., 29., 59.])          # identity deltas give the box back
    w1 = np.exp(np.float32(orc.BBOX_XFORM_CLIP)) * 20                              # dw clamped to log(1000/16)
    assert abs((out[0, 6] - out[0, 4] + 1).item() - w1) < 1e-2
    assert abs(((out[0, 5] + out[0, 7] + 1) / 2).item() - (40. + (-0.2) * 40)) < 1e-4


def test_roi_count_limits_rows():
    sizes = [(200, 300), (200, 300)]
    lg, rg, pr = orc.synth_box_head_outputs(2, 50, sizes, 5)
    full = orc.box_candidates(lg, rg, pr, sizes, orc.BoxPostParams())
    part = orc.box_candidates(lg, rg, pr, sizes, orc.BoxPostParams(), roi_count=[50, 17])
    np.testing.assert_array_equal(full[0][0], part[0][0])
    assert part[1][0].shape[0] == 17
    np.testing.assert_array_equal(full[1][0][:17], part[1][0])


def test_box_decode_against_the_reference_known_answer(golden_dir):
    """tests/test_box_coder.py of the reference (its own KAT, atol 1e-4), recorded by running it against the reference's
    BoxCoder: the restatement must equal what the reference computed bit for bit and satisfy the test's expectation."""
    import json

    with open(os.path.join(golden_dir, "box_coder_kat.json")) as f:
        cases = json.load(f)["cases"]
    assert cases
    for c in cases:
        out = orc.box_decode(torch.tensor(c["deltas"], dtype=torch.float32), torch.tensor(c["boxes"], dtype=torch.float32),
                             tuple(c["weights"])).numpy()
        np.testing.assert_array_equal(out, np.asarray(c["decoded"], dtype=np.float32))
        np.testing.assert_allclose(out, np.asarray(c["expected_by_the_test"], dtype=np.float32), atol=1e-4)
