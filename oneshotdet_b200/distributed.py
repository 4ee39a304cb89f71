"""Episode-level data parallelism: episodes are independent, so rank r of G owns a contiguous
block of the batch and no data-path collective is needed; the only exchange is one fixed-shape all-gather of the
detections (NCCL over NVLink on the GPU box, gloo in the CPU tests), replacing the reference's pickle ->
ByteTensor -> two all_gathers -> unpickle (maskrcnn_benchmark/utils/comm.py:48-88,
engine/inference.py:133-152)."""
from __future__ import annotations

import ctypes

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
    """Double-buffered, asynchronous gather of the detections: every step's results are snapshotted into a packed
    payload [E_local, K + 1, 6] (rows 0..K-1: x1, y1, x2, y2, score, global episode id; row K: the count in column 0);
    ``steps_per_gather`` consecutive payloads form one group, and ONE all_gather_into_tensor per group is issued with
    async_op=True, so NCCL moves group g over NVLink while the kernels of group g+1 already run.  The reference gathers
    once, after the whole dataset (engine/inference.py:133-152); ``steps_per_gather=1`` gathers every step (lowest
    latency to the consumer), larger groups amortise the collective's launch cost (scaling measurements in DESIGN
    section 7).  ``submit`` waits for the gather issued two groups earlier before reusing its buffers; ``finish`` issues
    the gather of a partially filled group and drains everything."""

    def __init__(self, e_local: int, k: int, device, episode_offset: int = 0, group=None, steps_per_gather: int = 1):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.e, self.k, self.m = e_local, k, max(1, int(steps_per_gather))
        self.payload, self.out, self.work = [], [], [None, None]
        for _ in range(2):
            p = torch.zeros((self.m, e_local, k + 1, 6), dtype=torch.float32, device=device)
            p[:, :, :k, 5] = torch.arange(episode_offset, episode_offset + e_local, device=device,
                                          dtype=torch.float32).view(1, -1, 1)
            self.payload.append(p)
            self.out.append(torch.empty((self.world * self.m, e_local, k + 1, 6), dtype=torch.float32, device=device))
        self.i = 0           # steps submitted
        self.filled = [0, 0]  # steps of the group currently held by each slot

    def _issue(self, slot):
        if self.world > 1:
            self.work[slot] = dist.all_gather_into_tensor(self.out[slot], self.payload[slot], group=self.group, async_op=True)
        else:
            self.out[slot].copy_(self.payload[slot])

    def submit(self, boxes, scores, count):
        """Snapshot one step; returns (slot, row) of the group buffer that will hold it after the gather."""
        g, row = divmod(self.i, self.m)
        slot = g & 1
        self.i += 1
        if row == 0:
            if self.work[slot] is not None:   # the gather issued two groups ago still reads this slot's payload
                self.work[slot].wait()
                self.work[slot] = None
            self.filled[slot] = 0
        p = self.payload[slot][row]
        p[:, :self.k, :4].copy_(boxes)
        p[:, :self.k, 4].copy_(scores)
        p[:, self.k, 0].copy_(count)
        self.filled[slot] = row + 1
        if row == self.m - 1:
            self._issue(slot)
        return slot, row

    def finish(self):
        g, row = divmod(self.i, self.m)
        if row != 0:                  # a partially filled group: gather what is there (stale rows are ignored by result())
            self._issue(g & 1)
            self.i = (g + 1) * self.m
        for j, w in enumerate(self.work):
            if w is not None:
                w.wait()
        self.work = [None, None]

    def result(self, slot, row: int = 0):  # noqa: D401  (packed-payload mode)
        """(dets [E_total, K, 6], counts int32 [E_total]) of step ``row`` of a finished group ``slot``; episodes are in
        rank order (rank r owns the contiguous block r of the batch, see shard_range)."""
        o = self.out[slot].view(self.world, self.m, self.e, self.k + 1, 6)[:, row].reshape(self.world * self.e, self.k + 1, 6)
        return o[:, :self.k], o[:, self.k, 0].to(torch.int32)


class BlockGatherer:
    """Copy-free variant for a double-buffered pipeline (``EpisodePipeline(double_buffer=True)``): the post-processing
    stage writes boxes | scores | index | count of a step into ONE result block (``FcosResult.block``), and that block is
    what goes over NVLink -- no packing kernels.  ``submit(block)`` issues one asynchronous ``all_gather_into_tensor`` of
    the block; before issuing it makes the current stream wait for the gather two steps back, whose source block the next
    step will overwrite -- call ``acquire()`` before launching a step and ``submit(block)`` after it.  ``result(slot)`` returns the (boxes, scores, index, count) views of every rank's block."""

    def __init__(self, e_local: int, k: int, device, group=None):
        from . import ops

        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.e, self.k = e_local, k
        _, self.block_bytes = ops.result_block_layout(e_local, k)
        self.out = [torch.empty((self.world * self.block_bytes,), dtype=torch.uint8, device=device) for _ in range(2)]
        self.work = [None, None]
        self.i = 0

    def acquire(self, n: int = 1):
        """Call BEFORE launching (or replaying) the step(s) whose result block(s) ``submit`` will ship next (``n``
        consecutive steps launched at once): makes the current
        stream wait for the gather issued two steps ago, which reads the block that step is about to overwrite.  (The
        wait inside ``submit`` alone comes too late -- by then the step's kernels are already enqueued.)"""
        for k in range(n):
            slot = (self.i + k) & 1
            if self.work[slot] is not None:
                self.work[slot].wait()
                self.work[slot] = None

    def submit(self, block):
        slot = self.i & 1
        self.i += 1
        if self.work[slot] is not None:     # acquire() was not called: still order the collective itself
            self.work[slot].wait()
            self.work[slot] = None
        if block.numel() != self.block_bytes:
            raise ValueError(f"result block of {block.numel()} bytes, expected {self.block_bytes}")
        if self.world > 1:
            self.work[slot] = dist.all_gather_into_tensor(self.out[slot], block, group=self.group, async_op=True)
        else:
            self.out[slot].copy_(block)
        return slot

    def finish(self):
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None, None]

    def result(self, slot):
        """List over ranks of (boxes [E,K,4], scores [E,K], index [E,K], count [E]) views of the gathered blocks."""
        from . import ops

        blocks = self.out[slot].view(self.world, self.block_bytes)
        return [ops.result_block(self.e, self.k, blocks.device, blocks[r])[1] for r in range(self.world)]


class PeerBlockGatherer:
    """The same exchange without a collective kernel: every rank owns a receive buffer [slots][world][block] that its
    peers map through CUDA IPC (handles travel once through the host-side process group); ``submit(block)`` PUSHES the
    step's result block into slot ``i % slots``, row ``rank`` of every rank's buffer -- one peer copy per destination on
    the copy engines, issued on a side stream behind an event of the compute stream (``mode='kernel'``: one small
    kernel of ours storing through the mapped peer pointers instead).  No SM is taken from the step's kernels and no
    rendezvous sits between the ranks, which is what NCCL's per-step all-gather kernel cost at 2-8 GPUs (DESIGN section
    7).  ``acquire()`` before a step makes the compute stream wait until the pushes that still read the block it
    overwrites have been issued to completion; ``finish()`` drains the pushes and runs a process-group barrier, after
    which ``result(slot)`` holds every rank's block of the last ``slots`` steps.  Between two ``finish`` (or ``fence``)
    calls at most ``slots`` steps may be submitted if a consumer reads the results on other ranks."""

    def __init__(self, e_local: int, k: int, device, group=None, slots: int = 2, mode: str = "copy"):
        import ctypes

        from . import _lib, ops

        if mode not in ("copy", "kernel"):
            raise ValueError("mode must be 'copy' or 'kernel'")
        self.group, self.mode, self.slots = group, mode, slots
        self.device = torch.device(device)
        ok = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ok else 1
        self.rank = dist.get_rank(group) if ok else 0
        self.e, self.k = e_local, k
        _, self.block_bytes = ops.result_block_layout(e_local, k)
        self.lib = _lib.load()
        _lib.require_device(self.device)
        nbytes = slots * self.world * self.block_bytes
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            _lib.check(self.lib.osd_comm_alloc(nbytes, ctypes.byref(p)), "osd_comm_alloc")
            self._own = p.value
            h = ctypes.create_string_buffer(64)
            _lib.check(self.lib.osd_comm_export(self._own, h), "osd_comm_export")
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, h.raw, group=group)
            else:
                handles[0] = h.raw
            self._peer = []
            for r, raw in enumerate(handles):
                if r == self.rank:
                    self._peer.append(self._own)
                    continue
                q = ctypes.c_void_p()
                _lib.check(self.lib.osd_comm_import(ctypes.create_string_buffer(raw, 64), ctypes.byref(q)), "osd_comm_import")
                self._peer.append(q.value)
        self._ptrs = (ctypes.c_void_p * self.world)()
        # receive buffer as a tensor view (no copy): [slots, world, block_bytes] uint8
        self.recv = _wrap_device_memory(self._own, nbytes, self.device).view(slots, self.world, self.block_bytes)
        self.side = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event() for _ in range(slots)]     # compute -> side: block written
        self._events = []                                           # side -> compute: pushes not yet waited for, oldest first
        self.i = 0

    def acquire(self, n: int = 1, in_flight: int | None = None):
        """Before launching the next ``n`` step(s): makes the current stream wait for the pushes that may still read a
        result block those steps overwrite -- every push issued so far except the ``in_flight`` most recent ones
        (default ``slots - n``: right for a pipeline that cycles through ``slots`` output sets in submission order;
        pass 0 when in doubt)."""
        keep = max(0, self.slots - n) if in_flight is None else max(0, int(in_flight))
        cur = torch.cuda.current_stream(self.device)
        while len(self._events) > keep:
            cur.wait_event(self._events.pop(0))

    def submit(self, block):
        import ctypes

        from . import _lib

        if block.numel() != self.block_bytes:
            raise ValueError(f"result block of {block.numel()} bytes, expected {self.block_bytes}")
        slot = self.i % self.slots
        self.i += 1
        off = (slot * self.world + self.rank) * self.block_bytes
        for r in range(self.world):
            self._ptrs[r] = self._peer[r] + off
        self.ready[slot].record(torch.cuda.current_stream(self.device))
        self.side.wait_event(self.ready[slot])
        fn = self.lib.osd_comm_push if self.mode == "copy" else self.lib.osd_comm_push_kernel
        with torch.cuda.device(self.device):
            _lib.check(fn(self._ptrs, self.world, block.data_ptr(), self.block_bytes, self.side.cuda_stream), "osd_comm_push")
        ev = torch.cuda.Event()
        ev.record(self.side)
        self._events.append(ev)
        if len(self._events) > 4 * self.slots:      # acquire() is not being called: do not grow without bound
            self._events.pop(0)
        return slot

    def push_on_current_stream(self, slot: int, block):
        """The same push as ``submit`` -- the block into row ``rank`` of receive slot ``slot`` of every rank -- issued on
        the CURRENT stream with no events and no bookkeeping.  Made for stream capture: recorded at the end of the CUDA
        graph of a post-processing chain (``EpisodePipeline.capture_streams(after_post=...)``) the exchange costs the host
        nothing per step, whatever the number of ranks (submit() issues world copies + two events from Python every step).
        The caller's stream order is the only ordering: a slot is rewritten by the stream that wrote it last."""
        from . import _lib

        if block.numel() != self.block_bytes:
            raise ValueError(f"result block of {block.numel()} bytes, expected {self.block_bytes}")
        if not 0 <= slot < self.slots:
            raise ValueError(f"slot {slot} out of range (0..{self.slots - 1})")
        ptrs = (ctypes.c_void_p * self.world)(*[self._peer[r] + (slot * self.world + self.rank) * self.block_bytes
                                               for r in range(self.world)])
        fn = self.lib.osd_comm_push if self.mode == "copy" else self.lib.osd_comm_push_kernel
        with torch.cuda.device(self.device):
            _lib.check(fn(ptrs, self.world, block.data_ptr(), self.block_bytes,
                          torch.cuda.current_stream(self.device).cuda_stream), "osd_comm_push")
        return slot

    def drain(self):
        """Device-side: the current stream waits for every push this rank has issued (no host synchronisation, no
        rendezvous).  What a timed region needs at its end; ``fence()`` is what a consumer of the received blocks needs."""
        cur = torch.cuda.current_stream(self.device)
        while self._events:
            cur.wait_event(self._events.pop(0))

    def fence(self):
        """All pushes issued so far by every rank have landed (process-group barrier behind a drained side stream)."""
        self.side.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    finish = fence

    def result(self, slot):
        """List over ranks of (boxes [E,K,4], scores [E,K], index [E,K], count [E]) views of the received blocks."""
        from . import ops

        return [ops.result_block(self.e, self.k, self.device, self.recv[slot, r])[1] for r in range(self.world)]

    def close(self):
        from . import _lib

        if getattr(self, "_own", None) is None:
            return
        with torch.cuda.device(self.device):
            self.side.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)       # nobody pushes into a buffer that is about to be unmapped
            for r, p in enumerate(self._peer):
                if r != self.rank:
                    _lib.check(self.lib.osd_comm_close(p), "osd_comm_close")
            self.recv = None
            _lib.check(self.lib.osd_comm_free(self._own), "osd_comm_free")
        self._own = None


def _wrap_device_memory(ptr: int, nbytes: int, device):
    """uint8 tensor view of raw device memory (``__cuda_array_interface__``); the caller keeps the memory alive."""

    class _Raw:
        pass

    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(raw, device=device)
