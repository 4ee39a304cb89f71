"""GPU parity of the second-stage ROI pooler (osd_roi_pool through the reference-shaped Pooler) against the oracle and
the executed-reference fixtures: fp32, bit-exact (every operation is a single rounded IEEE op in the reference's
order); the ROI -> level mapping goes through log2f and is compared as integers."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from test_oracle_pooler import POOLER_CASES, load_pooler

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALES = [1 / s for s in orc.FPN_STRIDES]


@pytest.mark.parametrize("name", POOLER_CASES)
def test_reference_fixtures(golden_dir, name):
    import oneshotdet_b200 as osd

    z, feats, boxes, scales, res, samp = load_pooler(golden_dir, name)
    sizes = [tuple(int(v) for v in hw) for hw in z["image_sizes"]]
    pooler = osd.Pooler((res, res), scales, samp)
    bl = [osd.BoxList(boxes[i].to(DEV), (sizes[i][1], sizes[i][0]), mode="xyxy") for i in range(boxes.size(0))]
    out = pooler([f.to(DEV) for f in feats], bl)
    assert tuple(out.shape) == tuple(z["pooled"].shape)
    np.testing.assert_array_equal(out.cpu().numpy(), z["pooled"])
    _, lv = osd.roi_pool([f.to(DEV) for f in feats], boxes.to(DEV), scales, res, samp, return_levels=True)
    np.testing.assert_array_equal(lv.cpu().numpy(), z["levels"].reshape(-1))


@pytest.mark.parametrize("res,samp,c", [(7, 2, 16), (1, 2, 8), (4, 0, 5), (14, 2, 3)])
def test_bit_exact_vs_oracle(res, samp, c):
    import oneshotdet_b200 as osd

    sizes = [(400, 600), (380, 640), (416, 500)]
    feats, _ = orc.synth_features(3, 1, c, 416, 640, seed=81 + res)
    rois = orc.synth_rois(3, 30, sizes, 82 + res)
    want, lw = orc.pooler_forward(feats, rois, SCALES, res, samp)
    got, lg = osd.roi_pool([f.to(DEV) for f in feats], rois.to(DEV), SCALES, res, samp, return_levels=True)
    np.testing.assert_array_equal(lg.cpu().numpy(), lw)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_roi_count_rows_are_zero_and_single_level_skips_the_mapper():
    import oneshotdet_b200 as osd

    sizes = [(256, 256), (256, 256)]
    feats, _ = orc.synth_features(2, 1, 4, 256, 256, seed=90)
    rois = orc.synth_rois(2, 20, sizes, 91)
    cnt = torch.tensor([20, 7], dtype=torch.int32, device=DEV)
    got, lv = osd.roi_pool([f.to(DEV) for f in feats], rois.to(DEV), SCALES, 7, 2, roi_count=cnt, return_levels=True)
    want, _ = orc.pooler_forward(feats, rois, SCALES, 7, 2)
    g = got.cpu().numpy()
    np.testing.assert_array_equal(g[:27], want[:27])
    assert not g[27:].any() and (lv.cpu().numpy()[27:] == -1).all()
    one, l1 = osd.roi_pool([feats[1].to(DEV)], rois.to(DEV), [SCALES[1]], 7, 2, return_levels=True)
    w1, _ = orc.pooler_forward([feats[1]], rois, [SCALES[1]], 7, 2)
    np.testing.assert_array_equal(one.cpu().numpy(), w1)
    assert not l1.any()


def test_full_size_properties_and_time():
    """Config-sized: 16 images x 2000 ROIs x 256 channels x 7x7 on the 800x1344 FPN maps.  Property: a constant feature
    map pools to that constant wherever all samples fall inside the map (interpolation weights sum to 1); spot rows are
    compared bit-exactly with the oracle."""
    import oneshotdet_b200 as osd

    b, r, c = 16, 2000, 256
    sizes = [(800, 1333)] * b
    shapes = orc.level_shapes(800, 1344)
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn((b, c, h, w), generator=g) for h, w in shapes]
    rois = orc.synth_rois(b, r, sizes, 95)
    dfe = [f.to(DEV) for f in feats]
    drois = rois.to(DEV)
    out = osd.roi_pool(dfe, drois, SCALES, 7, 2)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); out = osd.roi_pool(dfe, drois, SCALES, 7, 2); ev1.record(); torch.cuda.synchronize()
    print(f"\nroi_pool 16x2000x256x7x7 (channels-last copy + pool): {ev0.elapsed_time(ev1):.3f} ms, "
          f"{out.numel() * 4 / 1e9 / (ev0.elapsed_time(ev1) / 1e3):.0f} GB/s written")
    direct = osd.roi_pool(dfe, drois, SCALES, 7, 2, channels_last=False)
    ev0.record(); direct = osd.roi_pool(dfe, drois, SCALES, 7, 2, channels_last=False); ev1.record(); torch.cuda.synchronize()
    print(f"roi_pool direct NCHW taps: {ev0.elapsed_time(ev1):.3f} ms")
    assert torch.equal(out, direct)                                    # the two kernels produce the same bits
    del direct
    pick = [(0, 0), (3, 77), (15, 1999), (8, 1000)]
    for (i, j) in pick:                                                # each picked ROI against its own image
        want, _ = orc.pooler_forward([f[i:i + 1] for f in feats], rois[i:i + 1, j:j + 1], SCALES, 7, 2)
        np.testing.assert_array_equal(out[i * r + j].cpu().numpy(), want[0])
    ones = [torch.full_like(f, 1.5) for f in dfe]
    o1 = osd.roi_pool(ones, drois, SCALES, 7, 2)
    inside = (drois[..., 2] < 1300) & (drois[..., 3] < 780)
    v = o1.view(b, r, -1)[inside]
    assert torch.allclose(v, torch.full_like(v, 1.5), rtol=1e-6, atol=0)
