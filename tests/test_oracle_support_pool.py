"""CPU: pins the support-embedding oracle (oracle.roi_align_1x1) to the reference's own compiled ROIAlign
(maskrcnn_benchmark/csrc/cpu/ROIAlign_cpu.cpp, built unmodified into oracle/_ref)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc


@pytest.mark.parametrize("n,c,h,w,scale,ratio,seed", [(2, 8, 24, 24, 0.125, 2, 0), (3, 5, 12, 17, 0.0625, 2, 1),
                                                       (1, 4, 2, 2, 0.0078125, 2, 2), (2, 3, 6, 9, 0.03125, 0, 3),
                                                       (2, 6, 3, 3, 0.015625, 3, 4)])
def test_roi_align_1x1_equals_compiled_reference(ref_ops, n, c, h, w, scale, ratio, seed):
    if ref_ops is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.RandomState(seed)
    feat = rng.randn(n, c, h, w).astype(np.float32)
    # whole-image boxes [0, 0, size0, size1] like generalized_rcnn.py:257, plus jitter to leave the grid
    ext = np.stack((np.zeros(n), np.zeros(n), rng.uniform(0.5, 1.2, n) * w / scale, rng.uniform(0.5, 1.2, n) * h / scale), 1)
    ext[:, :2] = rng.uniform(-3, 3, (n, 2))
    rois = ext.astype(np.float32)
    rois5 = torch.cat((torch.arange(n, dtype=torch.float32)[:, None], torch.from_numpy(rois)), dim=1)
    ref = ref_ops.roi_align_forward(torch.from_numpy(feat), rois5, float(scale), 1, 1, int(ratio)).numpy()[:, :, 0, 0]
    got = orc.roi_align_1x1(feat, rois, scale, ratio)
    np.testing.assert_array_equal(got, ref)
