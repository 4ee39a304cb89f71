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
