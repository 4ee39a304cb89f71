"""EpisodePipeline: the public, fixed-shape entry point for serving a batch of (target, support) episodes --
matching on P3-P7 followed by the fused FCOS post-processing -- with optional episode sharding over the GPUs of
one box (one process per GPU, final all-gather of detections over NCCL).

    pipe = EpisodePipeline(batch=16, height=800, width=1344, image_sizes=[(800, 1333)] * 16)
    det = pipe.run_host(host_inputs)        # pinned host buffers in, host detections out  (end-to-end)
    res = pipe.run()                        # inputs already resident in pipe.features / pipe.cls ... (device)

The FCOS head itself (cuDNN convolutions between the two stages, modeling/rpn/fcos/fcos.py:12-99) is outside
the accelerated path: ``run`` consumes head outputs the caller provides."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops

FPN_STRIDES = (8, 16, 32, 64, 128)


def level_shapes(height: int, width: int, strides=FPN_STRIDES):
    return [(-(-height // s), -(-width // s)) for s in strides]


@dataclass
class PostParams:
    pre_nms_thresh: float = 0.0     # two-stage values of inference.py:337-351 / the shipped yaml :20-26
    pre_nms_top_n: int = 6000
    nms_thresh: float = 0.8
    fpn_post_nms_top_n: int = 2000
    min_size: float = 0.0


class EpisodePipeline:
    def __init__(self, batch: int, height: int, width: int, image_sizes, channels: int = 256, shots: int = 1,
                 strides=FPN_STRIDES, params: PostParams | None = None, match_mode: str = "product",
                 device="cuda", dtype=torch.float32, early_exit: bool = True, strict_iou: bool = False,
                 double_buffer: bool = False, pipeline_depth: int = 1):
        self.device = torch.device(device)
        self.batch, self.channels, self.shots = batch, channels, shots
        self.params = params or PostParams()
        self.shapes = level_shapes(height, width, strides)
        self.strides = tuple(strides)
        dev = self.device
        # device-resident inputs (the caller writes into these, or uses run_host)
        self.features = [torch.empty((batch, channels, h, w), dtype=dtype, device=dev) for h, w in self.shapes]
        self.supp = [torch.empty((batch * shots, channels, 1, 1), dtype=dtype, device=dev) for _ in self.shapes]
        self.cls = [torch.empty((batch, 1, h, w), dtype=torch.float32, device=dev) for h, w in self.shapes]
        self.reg = [torch.empty((batch, 4, h, w), dtype=torch.float32, device=dev) for h, w in self.shapes]
        self.ctr = [torch.empty((batch, 1, h, w), dtype=torch.float32, device=dev) for h, w in self.shapes]
        self.match = ops.PreparedMatch(self.features, self.supp, batch, match_mode)
        self.combined = self.match.outs
        p = self.params
        self.post = ops.PreparedFcos(self.cls, self.reg, self.ctr, self.strides, image_sizes, p.pre_nms_thresh,
                                     p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n, p.min_size, strict_iou,
                                     early_exit, private_workspace=True)
        # double_buffer: a second set of OUTPUT tensors (same inputs, same workspace -- steps are stream-ordered);
        # consecutive steps alternate between the two, so a consumer (the multi-GPU gather) can still read step i while
        # step i+1 runs
        # pipeline_depth > 1: that many post-processing chains may be IN FLIGHT at once (``run_overlapped(n)``: the
        # latency-bound chains of n consecutive batches run side by side while their matching launches stream back to
        # back), so every chain owns its outputs AND its workspace
        self.depth = max(1, int(pipeline_depth))
        self.posts = [self.post]
        for i in range(1, self.depth * (2 if double_buffer else 1)):
            # chains that run side by side (positions 0 .. depth-1 of a multi-step call) need their own workspace; the
            # second buffer set of a double-buffered pipeline is used by a later, stream-ordered call and shares them
            shared = self.posts[i % self.depth].result.workspace if i >= self.depth else None
            self.posts.append(ops.PreparedFcos(self.cls, self.reg, self.ctr, self.strides, image_sizes, p.pre_nms_thresh,
                                               p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n, p.min_size, strict_iou,
                                               early_exit, workspace=shared, private_workspace=shared is None))
        self._step = 0
        self._host_out = None
        self._streams = None
        self._graph = None

    # ---- device-resident step -------------------------------------------------------------------
    def run(self) -> ops.FcosResult:
        """One pass of the hot path over the resident batch: 1 matching launch + the post-processing launches,
        back to back on the current stream."""
        self.match()
        return self._next_post()()

    def _next_post(self):
        post = self.posts[self._step % len(self.posts)]
        self._step += 1
        return post

    def run_overlapped(self, n: int = 1):
        """The same work software-pipelined over streams: the HBM-bound matching stream and the latency-bound
        post-processing chain run concurrently (the post-processing streams have the higher priority so their small
        CTAs slot in as matching CTAs retire).  In a deployment the overlapping pair is matching of batch i+1 and
        post-processing of batch i (post-processing consumes the FCOS head's output of its own batch); here both
        stages read resident inputs, so the pairing inside one call is equivalent.  ``n`` > 1 (<= pipeline_depth)
        issues n consecutive steps at once: n matching launches back to back on the matching stream and the n chains
        side by side on n streams -- a chain is a string of small dependent kernels that takes longer than one
        matching launch, but two of them fit next to two matching launches.  The current stream waits for everything
        before the call returns control to later work.  Returns the step's FcosResult (n == 1) or a list of n."""
        if n > self.depth:
            raise ValueError(f"run_overlapped({n}) needs pipeline_depth >= {n}")
        if self._streams is None or len(self._streams) < 1 + n:
            lo, hi = torch.cuda.Stream.priority_range()
            self._streams = tuple([torch.cuda.Stream(self.device, priority=lo)] +
                                  [torch.cuda.Stream(self.device, priority=hi) for _ in range(max(n, self.depth))])
        s_match = self._streams[0]
        cur = torch.cuda.current_stream(self.device)
        s_match.wait_stream(cur)
        with torch.cuda.stream(s_match):
            for _ in range(n):
                self.match()
        out = []
        for k in range(n):
            s_post = self._streams[1 + k]
            s_post.wait_stream(cur)
            with torch.cuda.stream(s_post):
                out.append(self._next_post()())
            cur.wait_stream(s_post)
        cur.wait_stream(s_match)
        return out[0] if n == 1 else out

    def capture(self, overlapped: bool = True, steps_per_graph: int = 1):
        """Capture one step (all launches of the streams, with their fork/join dependencies) into a CUDA graph and
        return ``replay() -> FcosResult``: a serving loop then pays one graph launch per batch instead of ~10 kernel
        launches plus stream bookkeeping from Python.  Inputs and outputs are the pipeline's resident buffers.
        ``steps_per_graph`` = n > 1 (needs pipeline_depth >= n, overlapped): one graph holds n consecutive steps as
        ``run_overlapped(n)`` issues them and ``replay()`` returns the list of their n results."""
        n = int(steps_per_graph)
        if n > 1:
            if not overlapped:
                raise ValueError("steps_per_graph > 1 is the overlapped multi-step form")
            if len(self.posts) % n != 0:
                raise ValueError("the number of output sets must be a multiple of steps_per_graph")
            step = lambda: self.run_overlapped(n)   # noqa: E731
        else:
            step = self.run_overlapped if overlapped else self.run
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):   # warm up outside capture (lazy attribute setup, stream creation)
                step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graphs = []
        for g in range(len(self.posts) // n):   # one graph per output buffer set(s); replay alternates between them
            self._step = g * n
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                res = step()
            graphs.append((graph, res))
        self._graph = graphs
        self._step = 0
        state = {"i": 0}

        def replay():
            graph, res = graphs[state["i"] % len(graphs)]
            state["i"] += 1
            graph.replay()
            return res

        return replay

    def capture_streams(self, after_post=None):
        """Continuous software pipelining instead of one graph per step: returns a ``StreamedSteps`` (needs
        ``pipeline_depth >= 2``).  ``after_post(k, result)`` is called INSIDE the capture of output set k's chain, right
        after its post-processing launches: what it enqueues on the current stream (e.g. the multi-GPU push of the
        result block, ``PeerBlockGatherer.push_on_current_stream``) becomes part of that chain's CUDA graph."""
        return StreamedSteps(self, after_post)

    def input_tensors(self):
        return self.features + self.supp + self.cls + self.reg + self.ctr

    # ---- end-to-end step (host buffers in, host detections out) ----------------------------------
    def make_host_inputs(self, pinned: bool = True):
        return [torch.empty(t.shape, dtype=t.dtype, pin_memory=pinned) for t in self.input_tensors()]

    def run_host(self, host_inputs):
        """H2D copies of every input, the device step, D2H of (boxes, scores, count); returns host tensors after
        synchronising the stream.  Bytes moved are in ``h2d_bytes`` / ``d2h_bytes``."""
        for dst, src in zip(self.input_tensors(), host_inputs):
            dst.copy_(src, non_blocking=True)
        res = self.run()
        if self._host_out is None:
            self._host_out = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                                   for t in (res.boxes, res.scores, res.count))
        for dst, src in zip(self._host_out, (res.boxes, res.scores, res.count)):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host_out

    @property
    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.input_tensors())

    @property
    def d2h_bytes(self) -> int:
        r = self.post.result
        return sum(t.numel() * t.element_size() for t in (r.boxes, r.scores, r.count))

    # ---- multi-GPU: episodes are sharded, detections gathered --------------------------------------
    def pack_detections(self, res: ops.FcosResult, episode_offset: int = 0):
        """[B, K, 6] fp32 rows (x1, y1, x2, y2, score, global episode id) + counts: the fixed-shape payload that
        replaces the reference's pickled-BoxList all_gather (utils/comm.py:48-88)."""
        b, k = res.scores.shape
        eid = torch.arange(episode_offset, episode_offset + b, device=res.scores.device, dtype=torch.float32)
        return torch.cat((res.boxes, res.scores.unsqueeze(-1), eid.view(b, 1, 1).expand(b, k, 1)), dim=-1), res.count


class StreamedSteps:
    """Steps of an EpisodePipeline as THREE self-ordered streams with no join between steps: the matching launches run
    back to back on one (low-priority) stream, the post-processing chains of consecutive batches alternate between two
    more.  A chain is a string of a dozen small dependent kernels that takes LONGER than one matching launch when it
    shares the SMs with the matching stream; a graph that holds one whole step (``capture()``) makes every step wait for
    its chain.  Here two chains are in flight next to the matching stream at any time, so the step rate is
    max(matching, chain / 2).  In deployment the chain of batch i consumes the FCOS head's output of batch i, which
    depends on matching i: the same three-stream shape with one event per batch.

        steps = pipe.capture_streams(); steps.begin()
        for ...: res, stream = steps.step()     # res: the FcosResult this step writes, ready in `stream` order
        steps.join()                            # the current stream waits for everything issued so far

    Output sets rotate (``len(pipe.posts)``, a multiple of 2): a set is rewritten only by the stream that wrote it."""

    def __init__(self, pipe: "EpisodePipeline", after_post=None):
        if pipe.depth < 2 or len(pipe.posts) % pipe.depth != 0:
            raise ValueError("StreamedSteps needs an EpisodePipeline(pipeline_depth>=2): chains with their own workspaces")
        self.pipe = pipe
        self.chains = pipe.depth                      # chains in flight = streams they alternate between
        dev = pipe.device
        lo, hi = torch.cuda.Stream.priority_range()
        self.s_match = torch.cuda.Stream(dev, priority=lo)
        self.s_post = [torch.cuda.Stream(dev, priority=hi) for _ in range(self.chains)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up outside capture (kernel attributes, lazy allocations)
            pipe.match()
            for post in pipe.posts:
                post()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g_match = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_match, stream=side):
            pipe.match()
        self.g_post, self.results = [], []
        for k, post in enumerate(pipe.posts):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                res = post()
                if after_post is not None:
                    after_post(k, res)
            self.g_post.append(g)
            self.results.append(res)
        self.i = 0

    def begin(self):
        cur = torch.cuda.current_stream(self.pipe.device)
        for s in [self.s_match] + self.s_post:
            s.wait_stream(cur)

    def next_stream(self):
        """The stream the NEXT step's chain runs on (where an exchange of its result has to be ordered)."""
        return self.s_post[self.i % self.chains]

    def step(self):
        k = self.i % len(self.g_post)
        s = self.s_post[self.i % self.chains]
        with torch.cuda.stream(self.s_match):
            self.g_match.replay()
        with torch.cuda.stream(s):
            self.g_post[k].replay()
        self.i += 1
        return self.results[k], s

    def join(self):
        cur = torch.cuda.current_stream(self.pipe.device)
        for s in [self.s_match] + self.s_post:
            cur.wait_stream(s)


class HostStreamer:
    """End-to-end serving loop over HOST batches, software-pipelined across steps: two EpisodePipelines (two sets of
    device inputs and outputs) alternate; batch i+1 crosses PCIe on a copy stream while batch i computes (CUDA-graph
    replay on the compute stream) and the detections of batch i-1 return on a third stream.  ``submit(host_inputs)``
    enqueues one batch and returns the host detections of the PREVIOUS batch (``None`` for the first call) -- one host
    synchronisation per returned result, on an event, never on the device; ``drain()`` returns the last one.  The
    per-step cost is max(H2D, compute, D2H) instead of their sum (``run_host``)."""

    def __init__(self, pipes, use_graph: bool = True):
        assert len(pipes) >= 2, "HostStreamer alternates between at least two pipelines"
        self.pipes = list(pipes)
        dev = self.pipes[0].device
        self.device = dev
        self.s_h2d, self.s_run, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
        n = len(self.pipes)
        self.ev_in = [torch.cuda.Event() for _ in range(n)]      # inputs of slot landed
        self.ev_run = [torch.cuda.Event() for _ in range(n)]     # compute of slot finished (inputs may be overwritten)
        self.ev_out = [torch.cuda.Event() for _ in range(n)]     # detections of slot are on the host
        self.host_out = [None] * n
        self.steps = []
        for p in self.pipes:
            if use_graph:
                with torch.cuda.stream(self.s_run):
                    self.steps.append(p.capture(overlapped=False))
            else:
                self.steps.append(p.run)
        torch.cuda.synchronize(dev)
        self.i = 0
        self._pending = None

    def submit(self, host_inputs):
        slot = self.i % len(self.pipes)
        pipe = self.pipes[slot]
        if self.i >= len(self.pipes):
            self.s_h2d.wait_event(self.ev_run[slot])     # the step that last read these device inputs has finished
        with torch.cuda.stream(self.s_h2d):
            for dst, src in zip(pipe.input_tensors(), host_inputs):
                dst.copy_(src, non_blocking=True)
            self.ev_in[slot].record(self.s_h2d)
        self.s_run.wait_event(self.ev_in[slot])
        if self.i >= len(self.pipes):
            self.s_run.wait_event(self.ev_out[slot])     # its previous detections have left the device
        with torch.cuda.stream(self.s_run):
            res = self.steps[slot]()
            self.ev_run[slot].record(self.s_run)
        self.s_d2h.wait_event(self.ev_run[slot])
        with torch.cuda.stream(self.s_d2h):
            if self.host_out[slot] is None:
                self.host_out[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                                            for t in (res.boxes, res.scores, res.count))
            for dst, src in zip(self.host_out[slot], (res.boxes, res.scores, res.count)):
                dst.copy_(src, non_blocking=True)
            self.ev_out[slot].record(self.s_d2h)
        prev, self._pending = self._pending, slot
        self.i += 1
        if prev is None:
            return None
        self.ev_out[prev].synchronize()
        return self.host_out[prev]

    def drain(self):
        if self._pending is None:
            return None
        slot, self._pending = self._pending, None
        self.ev_out[slot].synchronize()
        return self.host_out[slot]
