"""CPU: pins the NMS oracle (oracle/nms_oracle.c) to the reference's golden vectors
(/root/reference/tests/test_nms.py) and to the reference's own compiled nms_cpu."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from helpers import clustered_boxes, random_boxes


def _kat(golden_dir):
    with open(os.path.join(golden_dir, "nms_kat.json")) as f:
        return json.load(f)["cases"]


def test_oracle_matches_reference_known_answers(golden_dir):
    cases = _kat(golden_dir)
    assert len(cases) == 6  # 5 thresholds on the 5-box vector + the 53-box vector
    for c in cases:
        keep = orc.nms(np.asarray(c["boxes"], np.float32), np.asarray(c["scores"], np.float32), c["thresh"])
        assert keep.dtype == np.int64
        np.testing.assert_array_equal(keep, np.asarray(c["keep_sorted"]))  # ascending original index
    # the literal expectations of tests/test_nms.py:55 for the 5-box vector
    expect = [[1, 3], [1, 3], [1, 3], [1, 2, 3, 4], [0, 1, 2, 3, 4]]
    five = [c for c in cases if len(c["boxes"]) == 5]
    big = [c for c in cases if len(c["boxes"]) == 53]
    assert [round(c["thresh"], 3) for c in five] == [0.1, 0.3, 0.5, 0.8, 0.9]
    for c, e in zip(five, expect):
        assert c["keep_sorted"] == e
    assert len(big) == 1 and len(big[0]["keep_sorted"]) == 26 and big[0]["thresh"] == 0.5


def test_oracle_empty_and_single():
    assert orc.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)
    np.testing.assert_array_equal(orc.nms(np.array([[0, 0, 10, 10]], np.float32), np.array([0.3], np.float32), 0.5), [0])


def test_oracle_ge_vs_strict():
    # two identical-size boxes with IoU exactly 0.5: 10x10 (+1 convention) shifted
    # box a = [0,0,9,19] (10x20=200), b = [0,10,9,29] (200): inter = 10*10 = 100, union = 300 -> 1/3
    a = np.array([[0, 0, 9, 19], [0, 10, 9, 29]], np.float32)
    s = np.array([0.9, 0.8], np.float32)
    thr = np.float32(100.0) / np.float32(300.0)
    np.testing.assert_array_equal(orc.nms(a, s, float(thr)), [0])             # >= suppresses (nms_cpu.cpp:60)
    np.testing.assert_array_equal(orc.nms(a, s, float(thr), strict=True), [0, 1])  # > does not (nms.cu:60)


@pytest.mark.parametrize("n,thr,seed", [(1, 0.5, 0), (17, 0.3, 1), (64, 0.5, 2), (65, 0.7, 3), (1000, 0.6, 4),
                                       (3350, 0.6, 5), (2500, 0.8, 6)])
def test_oracle_equals_compiled_reference(ref_ops, n, thr, seed):
    if ref_ops is None:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    import torch

    rng = np.random.RandomState(seed)
    boxes, scores = (clustered_boxes if seed % 2 else random_boxes)(rng, n)
    ref = ref_ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
    got = orc.nms(boxes, scores, thr)
    np.testing.assert_array_equal(got, ref)
    assert 0 < got.shape[0] <= n


def test_oracle_ties_follow_given_order(ref_ops):
    """With tied scores the reference's result depends on ATen's unstable sort; the oracle
    reproduces it exactly when handed that order, which pins the tie semantics."""
    if ref_ops is None:
        pytest.skip("oracle/_ref not built")
    import torch

    rng = np.random.RandomState(7)
    boxes, scores = random_boxes(rng, 2000, distinct_scores=False)
    order = torch.from_numpy(scores).sort(0, descending=True)[1].numpy()
    ref = ref_ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
    np.testing.assert_array_equal(orc.nms(boxes, scores, 0.5, order=order), ref)


def test_batched_oracle_is_segmentwise():
    rng = np.random.RandomState(9)
    sizes = [0, 5, 130, 1, 64, 0, 257]
    seg = np.concatenate(([0], np.cumsum(sizes)))
    boxes, scores = clustered_boxes(rng, int(seg[-1]), clusters=5)
    keep, counts = orc.batched_nms(boxes, scores, seg, 0.5)
    pos = 0
    for e, n in enumerate(sizes):
        lo = seg[e]
        k = orc.nms(boxes[lo:lo + n], scores[lo:lo + n], 0.5) + lo
        np.testing.assert_array_equal(keep[pos:pos + counts[e]], k)
        pos += counts[e]
    assert pos == keep.shape[0]
