"""GPU parity: libosd_b200's NMS (through the reference-shaped ``nms`` / ``boxlist_nms`` and the batched C-ABI
call) against the oracle -- bit-exact keep indices and counts."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from helpers import clustered_boxes, random_boxes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_nms(boxes, scores, thr, strict=False):
    import oneshotdet_b200 as osd

    keep = osd.ops.nms(torch.from_numpy(boxes).to(DEV), torch.from_numpy(scores).to(DEV), thr, strict=strict)
    return keep.cpu().numpy()


def test_known_answer_vectors(golden_dir):
    """The reference's own tests/test_nms.py vectors, through the drop-in `layers.nms`."""
    from oneshotdet_b200.layers import nms as box_nms

    with open(os.path.join(golden_dir, "nms_kat.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        boxes = torch.tensor(c["boxes"], dtype=torch.float32, device=DEV)
        scores = torch.tensor(c["scores"], dtype=torch.float32, device=DEV)
        keep = box_nms(boxes, scores, c["thresh"])
        assert keep.dtype == torch.int64 and keep.is_cuda
        np.testing.assert_array_equal(keep.cpu().numpy(), np.asarray(c["keep_sorted"]))


def test_empty_input_contract():
    from oneshotdet_b200.layers import nms as box_nms

    keep = box_nms(torch.zeros((0, 4), device=DEV), torch.zeros((0,), device=DEV), 0.5)
    assert keep.dtype == torch.int64 and keep.numel() == 0 and keep.device.type == "cpu"  # csrc/nms.h:17-18


@pytest.mark.parametrize("n,thr,seed,kind", [
    (1, 0.5, 0, "rand"), (2, 0.5, 1, "clu"), (63, 0.3, 2, "clu"), (64, 0.5, 3, "clu"), (65, 0.7, 4, "clu"),
    (129, 0.5, 5, "rand"), (1000, 0.6, 6, "clu"), (3350, 0.6, 7, "clu"), (4097, 0.8, 8, "clu"),
    (11600, 0.8, 9, "clu"), (11600, 0.5, 10, "rand"),
])
def test_bit_exact_vs_oracle(n, thr, seed, kind):
    rng = np.random.RandomState(seed)
    if kind == "clu":
        boxes, scores = clustered_boxes(rng, n, clusters=max(2, n // 60))
    else:
        boxes, scores = random_boxes(rng, n, extent=1200.0)
    ref = orc.nms(boxes, scores, thr)
    got = gpu_nms(boxes, scores, thr)
    assert got.dtype == np.int64
    np.testing.assert_array_equal(got, ref)
    assert 0 < ref.shape[0] < n or n <= 2


def test_strict_mode_matches_cuda_reference_semantics():
    a = np.array([[0, 0, 9, 19], [0, 10, 9, 29]], np.float32)  # IoU exactly 100/300
    s = np.array([0.9, 0.8], np.float32)
    thr = float(np.float32(100.0) / np.float32(300.0))
    np.testing.assert_array_equal(gpu_nms(a, s, thr), [0])
    np.testing.assert_array_equal(gpu_nms(a, s, thr, strict=True), [0, 1])
    rng = np.random.RandomState(3)
    boxes, scores = clustered_boxes(rng, 2000, clusters=30)
    np.testing.assert_array_equal(gpu_nms(boxes, scores, 0.5, strict=True), orc.nms(boxes, scores, 0.5, strict=True))


def test_boundary_iou_values_take_the_exact_division():
    """Integer boxes give many pairs whose IoU equals simple fractions exactly; thresholds sitting on those values
    exercise the near-threshold band where the kernel falls back to the IEEE division."""
    rng = np.random.RandomState(11)
    n = 1500
    x1 = rng.randint(0, 60, n).astype(np.float32)
    y1 = rng.randint(0, 60, n).astype(np.float32)
    w = rng.randint(4, 24, n).astype(np.float32)
    h = rng.randint(4, 24, n).astype(np.float32)
    boxes = np.stack((x1, y1, x1 + w - 1, y1 + h - 1), 1).astype(np.float32)
    scores = ((rng.permutation(n) + 1.0) / (n + 1.0)).astype(np.float32)
    for thr in (0.5, 1.0 / 3.0, 0.25, 0.6, 2.0 / 3.0, 0.75):
        for strict in (False, True):
            np.testing.assert_array_equal(gpu_nms(boxes, scores, thr, strict), orc.nms(boxes, scores, thr, strict=strict))


def test_tied_scores_use_stable_order():
    rng = np.random.RandomState(12)
    boxes, scores = random_boxes(rng, 4000, distinct_scores=False)
    assert np.unique(scores).shape[0] < 100
    np.testing.assert_array_equal(gpu_nms(boxes, scores, 0.5), orc.nms(boxes, scores, 0.5))


def test_irregular_boxes_follow_the_reference_arithmetic():
    """Inverted / zero-area / huge boxes: the episode takes the all-division path; still identical to nms_cpu."""
    rng = np.random.RandomState(13)
    boxes, scores = random_boxes(rng, 700)
    boxes[::7, [0, 2]] = boxes[::7, [2, 0]]        # x2 < x1
    boxes[3::11, 3] = boxes[3::11, 1] - 1.0        # height 0 under the +1 convention
    boxes[5::13] *= 1e6
    for thr in (0.3, 0.7):
        np.testing.assert_array_equal(gpu_nms(boxes, scores, thr), orc.nms(boxes, scores, thr))


def test_thresholds_outside_unit_interval():
    rng = np.random.RandomState(14)
    boxes, scores = clustered_boxes(rng, 500, clusters=8)
    for thr in (1.0, 1.5, 1e-6):
        np.testing.assert_array_equal(gpu_nms(boxes, scores, thr), orc.nms(boxes, scores, thr))


def test_large_n_rank_sort_path():
    rng = np.random.RandomState(15)
    n = 20000  # > 16384: global rank sort instead of the shared-memory bitonic network
    boxes, scores = clustered_boxes(rng, n, clusters=400, extent=4000.0)
    np.testing.assert_array_equal(gpu_nms(boxes, scores, 0.6), orc.nms(boxes, scores, 0.6))


def test_batched_ragged_segments():
    import oneshotdet_b200 as osd

    rng = np.random.RandomState(16)
    sizes = [0, 5, 130, 1, 64, 0, 257, 3350, 2, 1000]
    seg = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    boxes, scores = clustered_boxes(rng, int(seg[-1]), clusters=50)
    keep, counts = osd.batched_nms(torch.from_numpy(boxes).to(DEV), torch.from_numpy(scores).to(DEV),
                                   torch.from_numpy(seg).to(DEV), max(sizes), 0.6)
    keep, counts = keep.cpu().numpy(), counts.cpu().numpy()
    ref_keep, ref_counts = orc.batched_nms(boxes, scores, seg, 0.6)
    np.testing.assert_array_equal(counts, ref_counts)
    pos = 0
    for e, n in enumerate(sizes):
        np.testing.assert_array_equal(keep[seg[e]:seg[e] + counts[e]], ref_keep[pos:pos + ref_counts[e]])
        pos += ref_counts[e]


def test_boxlist_nms_dropin():
    import oneshotdet_b200 as osd

    rng = np.random.RandomState(17)
    boxes, scores = clustered_boxes(rng, 900, clusters=12)
    bl = osd.BoxList(torch.from_numpy(boxes).to(DEV), (1000, 1000), "xyxy")
    bl.add_field("scores", torch.from_numpy(scores).to(DEV))
    out = osd.boxlist_nms(bl, 0.5)
    ref = orc.nms(boxes, scores, 0.5)
    np.testing.assert_array_equal(out.bbox.cpu().numpy(), boxes[ref])
    np.testing.assert_array_equal(out.get_field("scores").cpu().numpy(), scores[ref])
    assert osd.boxlist_nms(bl, 0.0) is bl                       # boxlist_ops.py:22-23
    assert len(osd.boxlist_nms(bl, 0.5, max_proposals=7)) == 7  # boxlist_ops.py:31-32


def test_properties_at_full_size():
    """Size-independent checks at the BASELINE candidate count (11 600): ascending unique indices, idempotence
    (NMS of the survivors keeps all of them), and every dropped box overlaps a better kept box."""
    rng = np.random.RandomState(18)
    n = 11600
    boxes, scores = clustered_boxes(rng, n, clusters=300, extent=1300.0)
    keep = gpu_nms(boxes, scores, 0.8)
    assert np.all(np.diff(keep) > 0) and keep.min() >= 0 and keep.max() < n
    again = gpu_nms(boxes[keep], scores[keep], 0.8)
    np.testing.assert_array_equal(again, np.arange(keep.shape[0]))
    dropped = np.setdiff1d(np.arange(n), keep)[:200]
    kb, ks = torch.from_numpy(boxes[keep]), torch.from_numpy(scores[keep])
    for j in dropped:
        b = torch.from_numpy(boxes[j])
        better = ks > scores[j]
        xx1 = torch.maximum(kb[:, 0], b[0]); yy1 = torch.maximum(kb[:, 1], b[1])
        xx2 = torch.minimum(kb[:, 2], b[2]); yy2 = torch.minimum(kb[:, 3], b[3])
        inter = (xx2 - xx1 + 1).clamp(min=0) * (yy2 - yy1 + 1).clamp(min=0)
        area = (kb[:, 2] - kb[:, 0] + 1) * (kb[:, 3] - kb[:, 1] + 1)
        iou = inter / (area + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter)
        assert bool(((iou >= 0.8) & better).any())


def test_torch_extension_binding_matches_oracle(golden_dir):
    """`_C.nms` through the thin torch C++ extension (csrc/torch_ext.cpp): same kernels, same keeps."""
    from oneshotdet_b200 import _C

    ext = pytest.importorskip("oneshotdet_b200._C_torch")
    assert _C.BINDING == "torch-extension" and _C.nms is ext.nms
    with open(os.path.join(golden_dir, "nms_kat.json")) as f:
        for c in json.load(f)["cases"]:
            keep = _C.nms(torch.tensor(c["boxes"], dtype=torch.float32, device=DEV),
                          torch.tensor(c["scores"], dtype=torch.float32, device=DEV), c["thresh"])
            assert keep.dtype == torch.int64 and keep.is_cuda
            np.testing.assert_array_equal(keep.cpu().numpy(), np.asarray(c["keep_sorted"]))
    for n, thr, seed in [(1, 0.5, 0), (777, 0.6, 1), (6000, 0.8, 2)]:
        boxes, scores = clustered_boxes(np.random.RandomState(seed), n, clusters=max(2, n // 60))
        keep = _C.nms(torch.from_numpy(boxes).to(DEV), torch.from_numpy(scores).to(DEV), thr)
        np.testing.assert_array_equal(keep.cpu().numpy(), orc.nms(boxes, scores, thr))
    empty = _C.nms(torch.zeros((0, 4), device=DEV), torch.zeros((0,), device=DEV), 0.5)
    assert empty.dtype == torch.int64 and empty.numel() == 0 and empty.device.type == "cpu"
    with pytest.raises(RuntimeError):
        _C.nms(torch.zeros((3, 4)), torch.zeros((3,)), 0.5)          # no CPU path
