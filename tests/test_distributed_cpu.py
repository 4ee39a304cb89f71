"""CPU, world_size 2 over gloo: the episode sharding and the fixed-shape detection all-gather that replace the
reference's pickle all_gather (maskrcnn_benchmark/utils/comm.py:48-88, engine/inference.py:133-152)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oneshotdet_b200.distributed import gather_detections, shard_range, unpack_detections


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 16, 128, 129):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_eps, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total_eps, rank, world)
        e_local = -(-total_eps // world)           # padded so that every rank sends the same shape
        g = torch.Generator().manual_seed(1234)    # same stream on every rank: global ground truth
        all_boxes = torch.rand(total_eps, k, 4, generator=g)
        all_scores = torch.rand(total_eps, k, generator=g)
        all_counts = torch.randint(0, k + 1, (total_eps,), generator=g, dtype=torch.int32)
        dets = torch.full((e_local, k, 6), -1.0)
        counts = torch.zeros(e_local, dtype=torch.int32)
        n = hi - lo
        dets[:n, :, :4] = all_boxes[lo:hi]
        dets[:n, :, 4] = all_scores[lo:hi]
        dets[:n, :, 5] = torch.arange(lo, hi, dtype=torch.float32).view(n, 1)
        counts[:n] = all_counts[lo:hi]
        gd, gc = gather_detections(dets, counts)
        assert gd.shape == (world * e_local, k, 6) and gc.shape == (world * e_local,)
        out = unpack_detections(gd, gc, total_eps)
        ok = len(out) == total_eps
        for e, (b, s) in enumerate(out):
            c = int(all_counts[e])
            ok = ok and torch.equal(b, all_boxes[e, :c]) and torch.equal(s, all_scores[e, :c])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_eps", [8, 7])
def test_gather_detections_world_size_2(total_eps):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_eps, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_single_process_is_identity():
    d = torch.zeros(3, 4, 6)
    c = torch.zeros(3, dtype=torch.int32)
    gd, gc = gather_detections(d, c)
    assert gd is d and gc is c


def _gatherer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oneshotdet_b200.distributed import DetectionGatherer

        e, k = 3, 4
        ok = True

        def synth(step, r):
            gen = torch.Generator().manual_seed(100 * step + r)
            return (torch.rand(e, k, 4, generator=gen), torch.rand(e, k, generator=gen),
                    torch.randint(0, k + 1, (e,), generator=gen, dtype=torch.int32))

        # per-step gathers (5 steps > 2 buffers: exercises the reuse wait), then groups of 3 with a partial last group
        for m, steps in ((1, 5), (3, 8)):
            g = DetectionGatherer(e, k, torch.device("cpu"), episode_offset=rank * e, steps_per_gather=m)
            where = [g.submit(*synth(step, rank)) for step in range(steps)]
            g.finish()
            last_group = (steps - 1) // m
            for step in range(steps):
                if step // m < last_group - 1:
                    continue                     # that group's buffers have been reused since
                slot, row = where[step]
                assert (slot, row) == ((step // m) & 1, step % m)
                dets, counts = g.result(slot, row)
                for r in range(world):
                    boxes, scores, count = synth(step, r)
                    blk = dets[r * e:(r + 1) * e]
                    ok = ok and torch.equal(blk[..., :4], boxes) and torch.equal(blk[..., 4], scores)
                    ok = ok and torch.equal(counts[r * e:(r + 1) * e], count)
                    ok = ok and blk[:, 0, 5].tolist() == [float(r * e + i) for i in range(e)]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_async_detection_gatherer_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def _block_gatherer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oneshotdet_b200 import ops
        from oneshotdet_b200.distributed import BlockGatherer

        e, k = 3, 5
        g = BlockGatherer(e, k, torch.device("cpu"))
        blocks = [ops.result_block(e, k, torch.device("cpu")) for _ in range(2)]   # the double-buffered step outputs

        def fill(views, step, r):
            gen = torch.Generator().manual_seed(1000 * step + r)
            views[0].copy_(torch.rand(e, k, 4, generator=gen))
            views[1].copy_(torch.rand(e, k, generator=gen))
            views[2].copy_(torch.randint(0, 100, (e, k), generator=gen, dtype=torch.int32))
            views[3].copy_(torch.randint(0, k + 1, (e,), generator=gen, dtype=torch.int32))

        ok = True
        for step in range(5):
            block, views = blocks[step & 1]
            # the documented order: acquire() BEFORE the step overwrites the block the gather two steps back still reads
            g.acquire()
            assert g.work[step & 1] is None, "acquire() must have drained the gather that reads this block"
            fill(views, step, rank)
            slot = g.submit(block)
            assert slot == step & 1 and g.work[slot] is not None
        g.finish()
        for step in (3, 4):                                   # the two groups still resident
            per_rank = g.result(step & 1)
            assert len(per_rank) == world
            for r in range(world):
                _, want = ops.result_block(e, k, torch.device("cpu"))
                fill(want, step, r)
                ok = ok and all(torch.equal(a, b) for a, b in zip(per_rank[r], want))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_block_gatherer_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_block_gatherer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
