"""GPU: the public fixed-shape entry point (EpisodePipeline) -- serial, two-stream overlapped and host-buffer
(end-to-end) steps give identical detections, and those equal the oracle's."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(batch=3, h=256, w=320, c=32, shots=2):
    from oneshotdet_b200.pipeline import EpisodePipeline, PostParams

    sizes = [(250, 320), (256, 300), (256, 320)][:batch]
    p = PostParams(0.0, 400, 0.7, 100, 0.0)
    pipe = EpisodePipeline(batch, h, w, sizes, channels=c, shots=shots, params=p, device=DEV)
    feats, supp = orc.synth_features(batch, shots, c, h, w, seed=5)
    cls, reg, ctr = orc.synth_head_outputs(batch, h, w, seed=6)
    for dst, src in zip(pipe.input_tensors(), feats + supp + cls + reg + ctr):
        dst.copy_(src.to(DEV))
    return pipe, (feats, supp, cls, reg, ctr), sizes, orc.PostParams(0.0, 400, 0.7, 100, 0.0)


def snapshot(res):
    torch.cuda.synchronize()
    n = res.count.cpu().tolist()
    return n, [res.boxes[e, :k].cpu().clone() for e, k in enumerate(n)], [res.scores[e, :k].cpu().clone() for e, k in enumerate(n)]


def test_serial_overlapped_and_host_steps_agree():
    pipe, (feats, supp, cls, reg, ctr), sizes, p = build()
    a = snapshot(pipe.run())
    comb_a = [t.cpu().clone() for t in pipe.combined]
    b = snapshot(pipe.run_overlapped())
    comb_b = [t.cpu().clone() for t in pipe.combined]
    assert a[0] == b[0]
    for x, y in zip(a[1] + a[2] + comb_a, b[1] + b[2] + comb_b):
        assert torch.equal(x, y)
    host_in = pipe.make_host_inputs(pinned=True)
    for hbuf, d in zip(host_in, pipe.input_tensors()):
        hbuf.copy_(d)
    hb, hs, hc = pipe.run_host(host_in)
    assert hc.tolist() == a[0]
    for e, k in enumerate(a[0]):
        assert torch.equal(hb[e, :k], a[1][e]) and torch.equal(hs[e, :k], a[2][e])
    assert pipe.h2d_bytes == sum(t.numel() * t.element_size() for t in pipe.input_tensors())
    # matching equals the reference expression
    for o, e in zip(comb_a, orc.match_product(feats, supp, 3)):
        assert torch.equal(o, e)
    # detections: the oracle fed the device's own candidates returns the same boxes (NMS boundary)
    res = pipe.run()
    torch.cuda.synchronize()
    cb, cs, cl, cnt, slot = res.candidates()
    cb, cs, cnt = cb.cpu().numpy(), cs.cpu().numpy(), cnt.cpu().numpy()
    for e in range(3):
        gb = np.concatenate([cb[e, slot[l]:slot[l] + cnt[e, l]] for l in range(len(slot))])
        gs = np.concatenate([cs[e, slot[l]:slot[l] + cnt[e, l]] for l in range(len(slot))])
        eb, es, ek = orc.select_over_all_levels(gb, gs, p)
        assert a[0][e] == ek.shape[0]
        np.testing.assert_array_equal(a[1][e].numpy(), eb)


def test_pack_detections_layout():
    pipe, _, _, _ = build(batch=2)
    res = pipe.run()
    d, c = pipe.pack_detections(res, episode_offset=10)
    torch.cuda.synchronize()
    assert d.shape == (2, res.boxes.shape[1], 6) and c.dtype == torch.int32
    assert torch.equal(d[..., :4], res.boxes) and torch.equal(d[..., 4], res.scores)
    assert d[0, 0, 5].item() == 10.0 and d[1, 0, 5].item() == 11.0


def test_double_buffered_graph_steps_and_result_block():
    """double_buffer=True: consecutive steps alternate between two output sets (eager and CUDA-graph replay); each
    step's boxes | scores | index | count are views of ONE result block, which is what the multi-GPU gather ships."""
    from oneshotdet_b200 import ops
    from oneshotdet_b200.distributed import BlockGatherer
    from oneshotdet_b200.pipeline import EpisodePipeline, PostParams

    sizes = [(250, 320), (256, 300), (256, 320)]
    p = PostParams(0.0, 400, 0.7, 100, 0.0)
    pipe = EpisodePipeline(3, 256, 320, sizes, channels=32, shots=2, params=p, device=DEV, double_buffer=True)
    feats, supp = orc.synth_features(3, 2, 32, 256, 320, seed=5)
    cls, reg, ctr = orc.synth_head_outputs(3, 256, 320, seed=6)
    for dst, src in zip(pipe.input_tensors(), feats + supp + cls + reg + ctr):
        dst.copy_(src.to(DEV))
    r0 = pipe.run()
    r1 = pipe.run()
    assert r0 is not r1 and r0.block.data_ptr() != r1.block.data_ptr()
    a, b = snapshot(r0), snapshot(r1)
    assert a[0] == b[0] and all(torch.equal(x, y) for x, y in zip(a[1] + a[2], b[1] + b[2]))
    # the four outputs are views of the block
    _, views = ops.result_block(3, r0.boxes.size(1), DEV, r0.block)
    assert all(v.data_ptr() == t.data_ptr() for v, t in zip(views, (r0.boxes, r0.scores, r0.index, r0.count)))
    replay = pipe.capture(overlapped=True)
    g0, g1, g2 = replay(), replay(), replay()
    assert g0 is g2 and g0 is not g1
    c0, c1 = snapshot(g0), snapshot(g1)
    assert c0[0] == a[0] == c1[0] and all(torch.equal(x, y) for x, y in zip(a[1] + a[2], c1[1] + c1[2]))
    # single-process gatherer: the gathered block parses back to the same tensors
    g = BlockGatherer(3, r0.boxes.size(1), torch.device(DEV))
    slot = g.submit(g1.block)
    g.finish()
    (bx, sc, ix, ct), = g.result(slot)
    assert torch.equal(bx, g1.boxes) and torch.equal(sc, g1.scores) and torch.equal(ct, g1.count)


@pytest.mark.parametrize("double_buffer", [False, True])
def test_streamed_steps_match_the_serial_step(double_buffer):
    """capture_streams(): matching launches back to back on one stream, the post-processing chains of consecutive
    batches alternating between two more, no join between steps -- every step's detections and the matching output
    equal the serial step's, for several rotations of the output sets."""
    from oneshotdet_b200.pipeline import EpisodePipeline, PostParams

    sizes = [(250, 320), (256, 300), (256, 320)]
    p = PostParams(0.0, 400, 0.7, 100, 0.0)
    pipe = EpisodePipeline(3, 256, 320, sizes, channels=32, shots=2, params=p, device=DEV, double_buffer=double_buffer,
                           pipeline_depth=2)
    feats, supp = orc.synth_features(3, 2, 32, 256, 320, seed=5)
    cls, reg, ctr = orc.synth_head_outputs(3, 256, 320, seed=6)
    for dst, src in zip(pipe.input_tensors(), feats + supp + cls + reg + ctr):
        dst.copy_(src.to(DEV))
    ref = snapshot(pipe.run())
    comb = [t.cpu().clone() for t in pipe.combined]
    steps = pipe.capture_streams()
    assert len(steps.results) == (4 if double_buffer else 2)
    steps.begin()
    seen = []
    for i in range(9):
        res, stream = steps.step()
        seen.append((res, stream))
        assert stream is steps.s_post[i & 1] and res is steps.results[i % len(steps.results)]
    steps.join()
    torch.cuda.synchronize()
    for res in steps.results:
        got = snapshot(res)
        assert got[0] == ref[0] and all(torch.equal(x, y) for x, y in zip(got[1] + got[2], ref[1] + ref[2]))
    assert all(torch.equal(t.cpu(), c) for t, c in zip(pipe.combined, comb))
    with pytest.raises(ValueError):
        EpisodePipeline(3, 256, 320, sizes, channels=32, shots=2, params=p, device=DEV).capture_streams()
