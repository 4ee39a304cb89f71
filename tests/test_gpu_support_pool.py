"""GPU parity: the support-embedding producer (csrc/support_pool.cu) against the oracle; ROIAlign is bit-exact, the
average pool within 1e-6 relative (summation order)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALES = (0.125, 0.0625, 0.03125, 0.015625, 0.0078125)   # POOLER_SCALES of the shipped yaml (:29)


def support_maps(n, c, size, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(n, c, -(-size[0] // s), -(-size[1] // s), generator=g) for s in orc.FPN_STRIDES]


@pytest.mark.parametrize("n,c,size", [(2, 256, (192, 192)), (6, 64, (200, 320)), (1, 32, (64, 96))])
def test_supp_align_layer_bit_exact(n, c, size):
    import oneshotdet_b200 as osd
    from oneshotdet_b200.modeling.support_pooling import SuppAlignLayer

    feats = support_maps(n, c, size, seed=n + c)
    # the reference builds the box as [0, 0, size[0], size[1]] with size = (h, w) (generalized_rcnn.py:257)
    boxes = [osd.BoxList(torch.tensor([[0.0, 0.0, float(size[0]), float(size[1])]], device=DEV), size, "xyxy") for _ in range(n)]
    layer = SuppAlignLayer(SCALES, (1, 1), 2)
    out = layer([f.to(DEV) for f in feats], boxes)
    rois = np.tile(np.array([[0.0, 0.0, size[0], size[1]]], np.float32), (n, 1))
    ref = orc.support_pool_roialign(feats, rois, SCALES, 2)
    for o, r in zip(out, ref):
        assert o.shape == r.shape == (n, c, 1, 1)
        assert torch.equal(o.cpu(), r)


def test_adaptive_sampling_and_offset_boxes():
    from oneshotdet_b200.modeling.support_pooling import support_pool

    feats = support_maps(3, 16, (96, 128), seed=9)
    rng = np.random.RandomState(1)
    rois = np.stack((rng.uniform(-5, 20, 3), rng.uniform(-5, 20, 3), rng.uniform(60, 140, 3), rng.uniform(50, 110, 3)), 1).astype(np.float32)
    out = support_pool([f.to(DEV) for f in feats], torch.from_numpy(rois).to(DEV), SCALES, 0, "roialign")
    ref = orc.support_pool_roialign(feats, rois, SCALES, 0)
    for o, r in zip(out, ref):
        assert torch.equal(o.cpu(), r)


def test_avg_pool_and_feeds_matching():
    import oneshotdet_b200 as osd
    from oneshotdet_b200.modeling.support_pooling import SuppAvgPool

    b, s, c = 2, 2, 32
    supp_maps = support_maps(b * s, c, (96, 96), seed=4)
    pooled = SuppAvgPool()([f.to(DEV) for f in supp_maps])
    for o, r in zip(pooled, orc.support_pool_avg(supp_maps)):
        torch.testing.assert_close(o.cpu(), r, rtol=1e-5, atol=1e-6)
    # the pooled embeddings are what the matching module consumes
    feats, _ = orc.synth_features(b, s, c, 64, 96, seed=2)
    out = osd.MatchingModule("product", channels=c)([f.to(DEV) for f in feats], pooled, b)
    ref = orc.match_product(feats, [p.cpu() for p in pooled], b)
    for o, r in zip(out, ref):
        assert torch.equal(o.cpu(), r)
