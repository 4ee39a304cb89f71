"""Episode-level data parallelism: episodes are independent, so rank r of G owns a contiguous
block of the batch and no data-path collective is needed; the only exchange is one fixed-shape all-gather of the
detections (NCCL over NVLink on the GPU box, gloo in the CPU tests), replacing the reference's pickle ->
ByteTensor -> two all_gathers -> unpickle (maskrcnn_benchmark/utils/comm.py:48-88,
engine/inference.py:133-152)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_episodes: int, rank: int, world_size: int):
    """Contiguous block of episodes owned by `rank` (sizes differ by at most one; earlier ranks take the extra)."""
    base, extra = divmod(num_episodes, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, group=None):
    """dets [E_local, K, 6], counts int32 [E_local] -> ([E_total, K, 6], [E_total]) on every rank.
    E_local must be equal on all ranks (pad the last shard); one all_gather_into_tensor per tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dets, counts
    ws = dist.get_world_size(group)
    out_d = torch.empty((ws * dets.size(0),) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    out_c = torch.empty((ws * counts.size(0),), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_d, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(out_c, counts.contiguous(), group=group)
    return out_d, out_c


def unpack_detections(dets: torch.Tensor, counts: torch.Tensor, num_episodes: int | None = None):
    """Host-side compaction of a gathered payload: list of (boxes [n,4], scores [n]) per global episode id."""
    dets = dets.cpu()
    counts = counts.cpu().tolist()
    out = {}
    for e, n in enumerate(counts):
        if n == 0 and dets[e, 0, 5] < 0:
            continue  # padding episode
        gid = int(dets[e, 0, 5].item())
        out[gid] = (dets[e, :n, :4].clone(), dets[e, :n, 4].clone())
    ids = sorted(out) if num_episodes is None else range(num_episodes)
    return [out[i] for i in ids if i in out]


class DetectionGatherer:
    """Double-buffered, asynchronous gather of the per-step detections: the step's results are snapshotted into a
    packed payload [E_local, K + 1, 6] (rows 0..K-1: x1, y1, x2, y2, score, global episode id; row K: the count in
    column 0) and ONE all_gather_into_tensor is issued with async_op=True, so NCCL moves batch i over NVLink while
    the kernels of batch i+1 already run.  ``submit`` waits for the gather issued two steps earlier before reusing its
    buffers; ``finish`` drains everything."""

    def __init__(self, e_local: int, k: int, device, episode_offset: int = 0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.e, self.k = e_local, k
        self.payload, self.out, self.work = [], [], [None, None]
        for _ in range(2):
            p = torch.zeros((e_local, k + 1, 6), dtype=torch.float32, device=device)
            p[:, :k, 5] = torch.arange(episode_offset, episode_offset + e_local, device=device, dtype=torch.float32).view(-1, 1)
            self.payload.append(p)
            self.out.append(torch.empty((self.world * e_local, k + 1, 6), dtype=torch.float32, device=device))
        self.i = 0

    def submit(self, boxes, scores, count):
        slot = self.i & 1
        self.i += 1
        if self.work[slot] is not None:
            self.work[slot].wait()
        p = self.payload[slot]
        p[:, :self.k, :4].copy_(boxes)
        p[:, :self.k, 4].copy_(scores)
        p[:, self.k, 0].copy_(count)
        if self.world > 1:
            self.work[slot] = dist.all_gather_into_tensor(self.out[slot], p, group=self.group, async_op=True)
        else:
            self.out[slot].copy_(p)
        return slot

    def finish(self):
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None, None]

    def result(self, slot):
        """(dets [E_total, K, 6], counts int32 [E_total]) of a finished slot."""
        o = self.out[slot]
        return o[:, :self.k], o[:, self.k, 0].to(torch.int32)
