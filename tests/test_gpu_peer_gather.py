"""GPU test of the peer-memory detection exchange (csrc/peer_comm.cu, distributed.PeerBlockGatherer): two ranks, one
process each, gloo for the handle exchange, CUDA IPC for the data.  Both ranks may share one GPU (the driver's test
box has one): IPC between processes works on the same device as well.  Every rank checks that the blocks it received
equal the blocks the ranks sent, step by step, for the copy-engine push and for the store-kernel push."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
from oneshotdet_b200 import ops
from oneshotdet_b200.distributed import PeerBlockGatherer

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
dev = torch.device("cuda", rank % ngpu)
torch.cuda.set_device(dev)
dist.init_process_group("gloo")
E, K = 4, 50
for mode in ("copy", "kernel"):
    g = PeerBlockGatherer(E, K, dev, mode=mode)
    blocks = []
    for step in range(5):
        blk, (boxes, scores, index, count) = ops.result_block(E, K, dev)
        gen = torch.Generator(device=dev).manual_seed(1000 * step + rank)
        boxes.uniform_(0, 500, generator=gen); scores.uniform_(generator=gen)
        index.copy_(torch.arange(E * K, device=dev, dtype=torch.int32).view(E, K) + 7 * rank + step)
        count.fill_(rank + step)
        g.acquire()
        slot = g.submit(blk)
        blocks.append((slot, step, blk))
        if step % 2 == 1 or step == 4:
            g.fence()           # at most `slots` steps between fences when results are read on other ranks
            for s, st, b in blocks:
                res = g.result(s)
                assert len(res) == world
                assert torch.equal(g.recv[s, rank], b), (mode, st, "own block")
                for r in range(world):
                    rb, rs, ri, rc = res[r]
                    assert rc.tolist() == [r + st] * E, (mode, st, r, rc.tolist())
                    assert int(ri[0, 0]) == 7 * r + st, (mode, st, r)
            blocks = []
            dist.barrier()      # nobody pushes the next window before every rank has read this one
    g.close()
dist.barrier()
dist.destroy_process_group()
print("peer gather ok", rank)
"""


def test_peer_block_gatherer_two_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out = "TIMEOUT"
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "peer gather ok" in out, f"rank {r}:\n{out[-3000:]}"
