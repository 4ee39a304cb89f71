"""GPU parity through the thin torch C++ extension (csrc/torch_ext.cpp, module ``oneshotdet_b200._C_torch``): the
tensor-in / tensor-out forms of the matching module's forward and of FCOSPostProcessor.forward -- the two seams
besides ``nms`` that north_star asks the extension to keep (csrc/vision.cpp:7-15 is the module it stands in for)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ext():
    return pytest.importorskip("oneshotdet_b200._C_torch")


@pytest.mark.parametrize("mode", ["product", "concat", "concat_reversed"])
@pytest.mark.parametrize("b,s,c,h,w", [(2, 1, 64, 96, 160), (2, 3, 32, 100, 84)])
def test_match_forward_bit_exact(mode, b, s, c, h, w):
    m = ext()
    feats, supp = orc.synth_features(b, s, c, h, w, seed=11)
    out = m.match_forward([f.to(DEV) for f in feats], [x.to(DEV) for x in supp], b, mode)
    ref = {"product": orc.match_product, "concat": orc.match_concat,
           "concat_reversed": lambda f, x, n: orc.match_concat(f, x, n, reverse=True)}[mode](feats, supp, b)
    assert len(out) == len(ref)
    for o, r in zip(out, ref):
        assert o.is_cuda and o.shape == r.shape
        assert torch.equal(o.cpu(), r)


def test_match_forward_runs_on_the_current_stream_and_rejects_cpu():
    m = ext()
    feats, supp = orc.synth_features(1, 1, 32, 64, 64, seed=3)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = m.match_forward([f.to(DEV) for f in feats], [x.to(DEV) for x in supp], 1, "product")
    s.synchronize()
    for o, r in zip(out, orc.match_product(feats, supp, 1)):
        assert torch.equal(o.cpu(), r)
    with pytest.raises(RuntimeError):
        m.match_forward(feats, supp, 1, "product")           # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        m.match_forward([f.to(DEV) for f in feats], [x.to(DEV) for x in supp], 1, "fusion")


@pytest.mark.parametrize("params", [orc.PostParams(0.0, 500, 0.7, 120, 0.0), orc.PostParams(0.05, 300, 0.5, 50, 0.0)])
def test_fcos_postprocess_matches_oracle(params):
    from oneshotdet_b200 import _C

    m = ext()
    b, h, w = 2, 256, 320
    cls, reg, ctr = orc.synth_head_outputs(b, h, w, seed=77)
    sizes = [(250, 320), (256, 300)]
    p = params
    boxes, scores, index, count = _C.fcos_postprocess([t.to(DEV) for t in cls], [t.to(DEV) for t in reg],
                                                      [t.to(DEV) for t in ctr], sizes, orc.FPN_STRIDES, p.pre_nms_thresh,
                                                      p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n, p.min_size)
    assert _C.BINDING == "torch-extension" and m is not None
    ref = orc.fcos_postprocess(cls, reg, ctr, orc.FPN_STRIDES, sizes, p)
    cnt = count.cpu().numpy()
    for e in range(b):
        n = int(cnt[e])
        assert n == ref[e]["boxes"].shape[0]
        np.testing.assert_array_equal(boxes[e, :n].cpu().numpy(), ref[e]["boxes"])
        np.testing.assert_allclose(scores[e, :n].cpu().numpy(), ref[e]["scores"], rtol=2e-6, atol=0)
    assert index.dtype == torch.int32 and tuple(index.shape) == tuple(scores.shape)
    with pytest.raises(RuntimeError):
        m.fcos_postprocess(cls, reg, ctr, sizes, list(orc.FPN_STRIDES), 0.0, 100, 0.5, 10, 0.0)   # CPU tensors
