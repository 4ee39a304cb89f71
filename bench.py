#!/usr/bin/env python
"""bench.py -- matching + FCOS post-processing + NMS throughput (episodes/s) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A step = one pass of the hot path over one batch of synthetic episodes per GPU.  The headline workload is
BASELINE.json configs[1] ("config2": siamese FCOS R-50-FPN geometry, 16 episodes, 800x1333 target padded to 800x1344,
C=256, 1 shot, two-stage post-processing parameters 0 / 6000 / 0.8 / 2000): the product matching of P3-P7 (one
launch), then score -> per-level top-k -> decode/clip -> batched NMS -> post-NMS top-n.  The FCOS head between the two
stages is outside the path: head outputs are synthetic and resident.  With N > 1 every rank
owns its episodes (weak scaling) and every step's detections (the post-processing stage's result block) are pushed to
all ranks over NVLink peer memory on the copy engines (--gather peer; --gather block: one asynchronous NCCL all-gather
per step); the timed region ends when the launching stream has waited for every exchange this rank issued.  One NCCL
all-gather of the final step's block afterwards is the checker of the pushed copy.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric through the public
host-buffer API (H2D of all inputs + D2H of the detections inside the timed region), `roofline` describes the dominant
streaming kernel, `cpu_baseline` the reference's CPU path timed on this box's host cores on a bounded sample,
`workloads` the other BASELINE configs (config4, config5) and the NMS worst cases, each with a parity spot check.
`--impl reference` times the reference's own CPU path alone (rank 0 only under torchrun: `cpu_processes: 1`).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "matching+NMS episodes/s"
UNIT = "episodes/s"
CHANNELS = 256
TWO_STAGE = dict(pre_nms_thresh=0.0, pre_nms_top_n=6000, nms_thresh=0.8, fpn_post_nms_top_n=2000, min_size=0.0)
# single-stage branch of make_fcos_postprocessor (modeling/rpn/fcos/inference.py:326-336) with the stress values of
# BASELINE configs[4]; post-NMS cut = TEST.DETECTIONS_PER_IMG (config/defaults.py: 100)
STRESS = dict(pre_nms_thresh=0.01, pre_nms_top_n=1000, nms_thresh=0.6, fpn_post_nms_top_n=100, min_size=0.0)

WORKLOADS = {
    "config2": dict(batch=16, height=800, width=1344, image=(800, 1333), shots=1, params=TWO_STAGE, heads="spread",
                    early_exit=True,
                    text="configs[1]: siamese FCOS R-50-FPN geometry, 16 episodes/GPU, 800x1333 (padded 800x1344), C=256, "
                         "1 shot, fp32 product matching + score/top-k(6000)/decode/NMS(0.8)/top-2000"),
    "config4": dict(batch=64, height=1024, width=1024, image=(1024, 1024), shots=5, params=TWO_STAGE, heads="spread",
                    early_exit=True,
                    text="configs[3]: 64 episodes/GPU, 1024x1024 targets, 5-shot averaged support embedding, C=256, two-stage "
                         "post-processing parameters"),
    "config5": dict(batch=20, height=800, width=1344, image=(800, 1333), shots=1, params=STRESS, heads="spread",
                    early_exit=True,
                    text="configs[4] NMS stress: 20 episodes (classes) sharing one 800x1333 image size, score_thresh 0.01, "
                         "1000 pre-NMS candidates per level, NMS 0.6, 100 detections"),
    "nms_full": dict(batch=16, height=800, width=1344, image=(800, 1333), shots=1, params=TWO_STAGE, heads="clustered",
                     early_exit=True,
                     text="configs[1] geometry with clustered head outputs (one box shape per level -> long suppression "
                          "chains): the early exit misses, the second NMS pass and the dense sweep run"),
    "nms_noexit": dict(batch=16, height=800, width=1344, image=(800, 1333), shots=1, params=TWO_STAGE, heads="spread",
                       early_exit=False,
                       text="configs[1] inputs with early_exit=False: the full 11 600-candidate problem per episode "
                            "(67.3 M IoU pairs), what the CPU arm pays"),
}
LEVELS_TEXT = {(800, 1344): "P3-P7 100x168,50x84,25x42,13x21,7x11", (1024, 1024): "P3-P7 128x128,64x64,32x32,16x16,8x8"}


def config_dict(n_gpus, name="config2", gather=None):
    wl = WORKLOADS[name]
    return {"workload": wl["text"], "workload_key": name, "episodes_per_gpu": wl["batch"],
            "global_episodes": wl["batch"] * n_gpus, "levels": LEVELS_TEXT[(wl["height"], wl["width"])],
            "match_mode": "product", "match_dtype": "f32", "shots": wl["shots"], "post_params": wl["params"],
            "parallelism": f"episode-dp{n_gpus}", "gather": gather,
            "streams": "3 (matching | two alternating post-processing chains), see stages.overlap",
            "cache": "inputs larger than L2 (features in + out per step >= 734 MB vs 126 MB L2)"}


class StreamRunner:
    """K steps as three self-ordered streams (EpisodePipeline.capture_streams): matching launches back to back on one,
    the post-processing chains of consecutive batches alternating between two more; no join between steps.
    ``before(1)`` / ``after([res])`` are called in the stream context of the step's chain."""

    def __init__(self, pipe, after_post=None):
        self.pipe, self.depth, self.launch = pipe, 1, "cuda-graphs (1 matching + 1 chain per step) on 3 streams"
        self.steps = pipe.capture_streams(after_post)

    def run(self, k, before=None, after=None, mark=None):
        import torch

        self.steps.begin()
        last = None
        for _ in range(k):
            with torch.cuda.stream(self.steps.next_stream()):
                if before is not None:
                    before(1)
            last, s = self.steps.step()
            if after is not None:
                with torch.cuda.stream(s):
                    after([last])
        self.steps.join()
        if mark is not None:
            mark(k)
        return last


# ------------------------------------------------------------------------------------------------------
# synthetic inputs: features ~ N(0,1); cls ~ N(-4, 2^2); ctr ~ N(0,1); reg = exp(N(log 4s, .5^2));
# "clustered": reg = 12s exactly (identical box shapes per level, 24 strides wide: IoU with the neighbours one and two
# locations away is 0.92 / 0.85 > the 0.8 threshold -> heavy suppression, the early exit cannot hit)
# ------------------------------------------------------------------------------------------------------
def fill_inputs(pipe, seed, heads="spread"):
    import torch

    g = torch.Generator(device=pipe.device).manual_seed(seed)
    for t in pipe.features + pipe.supp:
        t.normal_(generator=g)
    for t in pipe.cls:
        t.normal_(mean=-4.0, std=2.0, generator=g)
    for t in pipe.ctr:
        t.normal_(generator=g)
    for t, s in zip(pipe.reg, pipe.strides):
        if heads == "clustered":
            t.fill_(12.0 * s)     # boxes of 24 strides: neighbours up to two locations away overlap above IoU 0.8
        else:
            t.normal_(mean=math.log(4.0 * s), std=0.5, generator=g).exp_()


def make_pipe(name, dev, double_buffer=False, depth=1):
    from oneshotdet_b200.pipeline import EpisodePipeline, PostParams

    wl = WORKLOADS[name]
    return EpisodePipeline(wl["batch"], wl["height"], wl["width"], [wl["image"]] * wl["batch"], channels=CHANNELS,
                           shots=wl["shots"], params=PostParams(**wl["params"]), match_mode="product", device=dev,
                           early_exit=wl["early_exit"], double_buffer=double_buffer, pipeline_depth=depth)


class StepRunner:
    """K steps of a pipeline as CUDA-graph replays: graphs of `depth` consecutive steps (their post-processing chains
    side by side, see EpisodePipeline.run_overlapped) plus a one-step graph for a remainder.  ``run(k, after=None)``
    executes exactly k steps; ``after(results)`` is called after every replay with the list of its results."""

    def __init__(self, pipe, depth, overlap=True, use_graph=True):
        import torch

        self.pipe, self.depth, self.launch = pipe, depth, "eager"
        self.one = pipe.run_overlapped if overlap else pipe.run
        self.many = (lambda: pipe.run_overlapped(depth)) if depth > 1 else None
        if use_graph:
            try:
                one = pipe.capture(overlapped=overlap)
                many = pipe.capture(overlapped=True, steps_per_graph=depth) if depth > 1 else None
                self.one, self.many, self.launch = one, many, "cuda-graph"
            except Exception as exc:  # noqa: BLE001  (launch mechanism only; the kernels are the same either way)
                print(f"[bench] CUDA-graph capture failed ({exc}); launching eagerly", file=sys.stderr)
                torch.cuda.synchronize()

    def run(self, k, before=None, after=None, mark=None):
        done, last = 0, None
        while done < k:
            n = self.depth if (self.many is not None and k - done >= self.depth) else 1
            if before is not None:
                before(n)
            res = self.many() if n > 1 else self.one()
            res = res if isinstance(res, list) else [res]
            if after is not None:
                after(res)
            done += n
            last = res[-1]
            if mark is not None:
                mark(done)
        return last


# ------------------------------------------------------------------------------------------------------
# the reference's CPU path on a bounded sample: torch.mul matching (generalized_rcnn.py:100-104, 306-311) + the
# reference's own FCOSPostProcessor.forward (modeling/rpn/fcos/inference.py:251-323, unmodified, vendored next to its
# compiled nms_cpu in oracle/_ref by oracle/build_ref.py).  Without the vendored copy: the oracle restatement.
# ------------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, name="config2", episodes=1, seed=7, threads=None, inputs=None):
        import torch

        from oracle import build_ref
        from oracle import oracle as orc

        self.orc, self.torch = orc, torch
        torch.set_num_threads(threads or os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        wl = WORKLOADS[name]
        self.wl, self.episodes = wl, episodes
        self.params = orc.PostParams(**wl["params"])
        self.sizes = [wl["image"]] * episodes
        if inputs is None:
            feats, supp = orc.synth_features(episodes, wl["shots"], CHANNELS, wl["height"], wl["width"], seed)
            cls, reg, ctr = orc.synth_head_outputs(episodes, wl["height"], wl["width"], seed, distinct=False)
            if wl["heads"] == "clustered":
                reg = [r * 0 + float(4 * s) for r, s in zip(reg, orc.FPN_STRIDES)]
        else:
            feats, supp, cls, reg, ctr = inputs
        self.feats, self.supp, self.cls, self.reg, self.ctr = feats, supp, cls, reg, ctr
        self.post = None
        self.kind = "port"
        self.nms_fn = None
        try:
            ref_c = build_ref.load_reference_python()
            if ref_c is not None:
                from maskrcnn_benchmark.modeling.rpn.fcos.fcos import FCOSModule  # noqa: PLC0415
                from maskrcnn_benchmark.modeling.rpn.fcos.inference import FCOSPostProcessor  # noqa: PLC0415

                cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(RPN_ONLY=False),
                                            FEW_SHOT=types.SimpleNamespace(ADD_ARTIFICIAL_PROPOSALS=False))
                p = self.params
                self.post = FCOSPostProcessor(cfg, p.pre_nms_thresh, p.pre_nms_top_n, p.nms_thresh, p.fpn_post_nms_top_n,
                                              p.min_size, num_classes=2, dense_points=1, score_calculator="BINARY").eval()
                fake = types.SimpleNamespace(dense_points=1)
                fake.get_dense_locations = lambda loc, stride, device: loc
                self.locations = [FCOSModule.compute_locations_per_level(fake, c.shape[-2], c.shape[-1], s, torch.device("cpu"))
                                  for c, s in zip(cls, orc.FPN_STRIDES)]
                self.kind = "reference"
        except Exception as exc:  # noqa: BLE001  (fall back to the restatement, and say so in `kind`)
            print(f"[bench] reference python unavailable ({exc}); timing the oracle restatement", file=sys.stderr)
            self.post = None
        if self.post is None:
            try:
                ref = build_ref.load_ref()
            except Exception:  # noqa: BLE001
                ref = None
            if ref is not None:
                self.kind = "port+reference-nms"
                self.nms_fn = lambda b, s, thr: ref.nms(torch.from_numpy(b), torch.from_numpy(s), float(thr)).numpy()

    def step(self, keep_result=False):
        orc, torch = self.orc, self.torch
        t0 = time.perf_counter()
        out = orc.match_product(self.feats, self.supp, self.episodes)   # the reference expression itself
        t1 = time.perf_counter()
        if self.post is not None:
            with torch.no_grad():
                res = self.post(self.locations, self.cls, self.reg, self.ctr, self.sizes)
        else:
            res = orc.fcos_postprocess(self.cls, self.reg, self.ctr, orc.FPN_STRIDES, self.sizes, self.params,
                                       nms_fn=self.nms_fn)
        t2 = time.perf_counter()
        assert len(out) == 5 and len(res) == self.episodes
        if keep_result:
            self.last = res
        return t1 - t0, t2 - t1

    def detections(self, e=0):
        """(boxes [n,4], scores [n]) of episode e of the last kept step, as numpy."""
        r = self.last[e]
        if self.post is not None:
            return r.bbox.numpy(), r.get_field("scores").numpy()
        return r["boxes"], r["scores"]

    def sample_text(self):
        if self.kind == "reference":
            post = ("the reference's own FCOSPostProcessor.forward (unmodified modeling/rpn/fcos/inference.py + "
                    "structures/boxlist_ops.py, vendored by oracle/build_ref.py) with its nms_cpu compiled unmodified")
        elif self.kind == "port+reference-nms":
            post = "FCOS post-processing restated with the reference's ATen ops + the reference's compiled nms_cpu"
        else:
            post = "FCOS post-processing restated with the reference's ATen ops + the C port of nms_cpu"
        return (f"{self.episodes} episode(s) of the workload per step: torch.mul matching on {self.cores} threads + {post}; "
                f"nms_cpu is single-threaded by construction (no OpenMP)")


def cpu_multi_process_throughput(processes=None, timed_steps=2):
    """The same CPU path run as `processes` independent single-threaded workers (one episode stream each) -- what the host
    cores deliver when the reference's per-image loop is data-parallelised over processes, since its nms_cpu cannot use a
    second thread.  Every worker is this script with --ref-worker; the aggregate rate is the sum of the workers' rates."""
    processes = processes or (os.cpu_count() or 1)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", str(100 + i), "--steps",
                               str(timed_steps)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env)
             for i in range(processes)]
    rate, ok = 0.0, 0
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=300)
            rate += float(out.strip().splitlines()[-1])
            ok += 1
        except Exception:  # noqa: BLE001  (a worker that died only lowers the reported rate)
            pr.kill()
    return {"value": rate, "unit": UNIT, "processes": ok, "threads_per_process": 1, "steps_per_worker": timed_steps,
            "what": "independent single-threaded workers, one episode per step each; sum of the workers' rates"}


def cpu_baseline_dict(single, kind, multi):
    """`value` = what ALL the host cores deliver on this path: the reference's nms_cpu cannot use a second thread, so
    the way to use every core is one single-threaded worker process per core, each running the unmodified reference on
    its own episodes (`multi`).  The reference run as it ships -- ONE process, torch.mul on all threads, NMS on one -- is
    reported beside it as `single_process`.  Without the multi-process measurement the single process is the value."""
    if multi is None or not multi.get("processes"):
        return {"value": single["value"], "unit": UNIT, "cores": single["cores"], "kind": kind, "sample": single["sample"],
                "single_process": single, "multi_process": None}
    return {"value": multi["value"], "unit": UNIT, "cores": multi["processes"] * multi["threads_per_process"], "kind": kind,
            "sample": f"{multi['processes']} worker processes x {multi['threads_per_process']} thread, each timing "
                      f"{multi.get('steps_per_worker', 2)} steps of 1 episode of the workload after one warm-up step: "
                      + re.sub(r" on \d+ threads", "", single["sample"].split(";")[0].split(": ", 1)[-1])
                      + "; sum of the workers' rates",
            "single_process": single, "multi_process": multi}


def run_ref_worker(seed, steps):
    ref = CpuReference(episodes=1, seed=seed, threads=1)
    ref.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step()
    print(steps * ref.episodes / (time.perf_counter() - t0), flush=True)
    return 0


def run_reference_arm(args):
    """The reference's CPU path with all the host threads it can use: one single-threaded worker process per core, each
    running the unmodified reference on its own episodes (its NMS cannot use a second thread); the one-process figure is
    in `cpu_baseline.single_process`.  Under torchrun (N > 1) rank 0 alone runs and prints the line, the other ranks
    exit without work (the contract of this tier): a per-N ratio against it compares N GPUs with this ONE host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(args.workload, episodes=1)
    for _ in range(max(args.warmup, 1)):
        ref.step()
    times = []
    budget_s = 150.0   # keep the whole run within a few minutes whatever K is; `steps` reports what was timed
    for _ in range(args.steps):
        a, b = ref.step()
        times.append(a + b)
        if sum(times) > budget_s:
            break
    nsteps = len(times)
    total = sum(times)
    single = {"value": ref.episodes * nsteps / total, "unit": UNIT, "cores": ref.cores, "processes": 1,
              "sample": ref.sample_text() + f"; {nsteps} steps, {total / nsteps:.3f} s/episode"}
    multi = None if args.no_multi_process else cpu_multi_process_throughput(timed_steps=max(2, min(args.steps, 8)))
    cpu = cpu_baseline_dict(single, ref.kind, multi)
    value = cpu["value"]
    procs = multi["processes"] if (multi and multi.get("processes")) else 1
    # a step of this arm = every worker process finishing one episode: `procs` episodes per step
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": (multi["steps_per_worker"] if procs > 1 else nsteps), "warmup": args.warmup,
            "ms_per_step": 1e3 * procs / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus, args.workload),
            "cpu_processes": procs,
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML polling thread)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        if self.ok:
            self._stop.set()
            self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram__bytes_read+write of the match kernel per launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "match_kernel_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def episode_inputs_to_host(pipe, e=0):
    """Episode e's inputs as CPU tensors in the layout the reference consumes."""
    s = pipe.shots
    feats = [t[e:e + 1].cpu() for t in pipe.features]
    supp = [t[e * s:(e + 1) * s].cpu() for t in pipe.supp]
    return feats, supp, [t[e:e + 1].cpu() for t in pipe.cls], [t[e:e + 1].cpu() for t in pipe.reg], \
        [t[e:e + 1].cpu() for t in pipe.ctr]


def parity_spot_check(name, pipe, res, with_cpu_time=True):
    """Episode 0 of the resident batch through the CPU reference path: detections must agree (boxes exactly, scores to
    2e-6 relative: the GPU sigmoid is 1/(1+expf(-x)) with IEEE division, ATen's is vectorised) and so must the matching
    output.  Returns the comparison and the CPU time for that episode."""
    import numpy as np
    import torch

    ref = CpuReference(name, episodes=1, inputs=episode_inputs_to_host(pipe, 0))
    tm, tp = ref.step(keep_result=True)
    rb, rs = ref.detections(0)
    n = int(res.count[0].item())
    gb = res.boxes[0, :n].cpu().numpy()
    gs = res.scores[0, :n].cpu().numpy()
    same_n = n == rb.shape[0]
    out = {"episode": 0, "checker": ref.kind, "count_gpu": n, "count_cpu": int(rb.shape[0]), "count_equal": bool(same_n)}
    if same_n and n > 0:
        out["boxes_equal"] = bool(np.array_equal(gb, rb))
        out["max_score_rel_err"] = float(np.max(np.abs(gs - rs) / np.maximum(np.abs(rs), 1e-30)))
        if not out["boxes_equal"]:
            # say what differs: rows, whether the two results hold the same boxes in another order, and the score gap
            # between the rows that trade places (a swap of two candidates whose scores differ by less than the stated
            # score tolerance is the sigmoid's rounding, not a different selection)
            rows = np.nonzero(np.any(gb != rb, axis=1))[0]
            order = lambda a: a[np.lexsort(a.T[::-1])]  # noqa: E731
            out["rows_differ"] = int(rows.size)
            out["first_rows_differ"] = rows[:8].tolist()
            out["same_boxes_other_order"] = bool(np.array_equal(order(gb), order(rb)))
            r0 = int(rows[0])
            out["first_diff"] = {"row": r0, "gpu_box": gb[r0].tolist(), "cpu_box": rb[r0].tolist(),
                                 "gpu_score": float(gs[r0]), "cpu_score": float(rs[r0]),
                                 "cpu_score_next": float(rs[min(r0 + 1, n - 1)])}
            # Exactly tied scores: the reference's order among them is whatever ATen's topk(sorted=False)
            # (fcos/inference.py:96-98) handed to its stable sorts -- unspecified; ours is "lower location first".  Rows that
            # differ only by a permutation inside a run of bit-identical CPU scores are the same result.
            runs_ok = True
            for r in rows.tolist():
                a = r
                while a > 0 and rs[a - 1] == rs[r]:
                    a -= 1
                b = r
                while b + 1 < n and rs[b + 1] == rs[r]:
                    b += 1
                runs_ok = runs_ok and b > a and np.array_equal(order(gb[a:b + 1]), order(rb[a:b + 1]))
            out["differ_only_inside_exact_score_ties"] = bool(runs_ok and out["same_boxes_other_order"])
    mref = ref.orc.match_product(ref.feats, ref.supp, 1)
    out["matching_equal"] = bool(all(torch.equal(o[0:1].cpu(), m) for o, m in zip(pipe.combined, mref)))
    boxes_ok = out.get("boxes_equal", n == 0) or out.get("differ_only_inside_exact_score_ties", False)
    out["ok"] = bool(same_n and boxes_ok and out.get("max_score_rel_err", 0.0) <= 2e-6 and out["matching_equal"])
    cpu = {"matching_s": tm, "post_s": tp, "episodes_per_s": 1.0 / (tm + tp), "kind": ref.kind, "cores": ref.cores} \
        if with_cpu_time else None
    return out, cpu


def time_steps(step_fn, steps, torch):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        step_fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[-1]) / steps, statistics.median(per)


def run_extra_workload(name, dev, steps, rank, use_graph=True, depth=1, streamed=False):
    """One of the non-headline workloads on this GPU: overlapped/graph step time, isolated stage times, parity spot
    check against the CPU reference path on episode 0, and that path's time on the same inputs."""
    import torch

    wl = WORKLOADS[name]
    pipe = make_pipe(name, dev, depth=2 if streamed else depth)
    fill_inputs(pipe, seed=3000 + 17 * rank + sum(map(ord, name)), heads=wl["heads"])
    runner = StreamRunner(pipe) if streamed else StepRunner(pipe, depth, overlap=True, use_graph=use_graph)
    launch = runner.launch
    runner.run(4 if streamed else 2 * depth)
    torch.cuda.synchronize()
    n = max(8, min(steps, 20)) if streamed else max(4, min(steps, 10)) // depth * depth
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    runner.run(n)
    ev[1].record()
    torch.cuda.synchronize()
    avg_ms = ev[0].elapsed_time(ev[1]) / n
    med_ms = avg_ms
    m_ms, _ = time_steps(pipe.match, n, torch)
    p_ms, _ = time_steps(pipe.post, n, torch)
    res = pipe.run()
    torch.cuda.synchronize()
    kept = res.kept_before_cut().cpu().tolist()
    parity, cpu = parity_spot_check(name, pipe, res)
    out = {"workload": wl["text"], "episodes_per_gpu": wl["batch"], "post_params": wl["params"], "shots": wl["shots"],
           "early_exit": wl["early_exit"], "ms_per_step": avg_ms, "steps_in_flight": depth,
           "episodes_per_s": wl["batch"] / (avg_ms * 1e-3), "steps": n, "launch": launch,
           "stages": {"match_ms_isolated": m_ms, "post_ms_isolated": p_ms},
           "detections_per_episode": res.count.cpu().tolist()[:4], "kept_before_cut": kept[:4],
           "parity": parity, "cpu_same_inputs": cpu}
    del pipe
    torch.cuda.empty_cache()
    return out


def h2d_ceiling(pipe, host_in, steps, torch, dist, world, dev):
    """What the box delivers for this step's host->device traffic alone: the same pinned buffers copied into the same
    device tensors back to back, all ranks at once, nothing else running."""
    s = torch.cuda.Stream(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(steps):
            for dst, src in zip(pipe.input_tensors(), host_in):
                dst.copy_(src, non_blocking=True)
    s.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return pipe.h2d_bytes * steps / float(t.item()) / 1e9


# ------------------------------------------------------------------------------------------------------
def run_second_stage(dev, steps):
    """The second stage at the first stage's output size -- 16 episodes x 2000 proposals, C = 256, MLP 1024 (SURVEY section
    8(f) row 2): multi-level ROI pooler (bf16 [roi, pixel, channel] rows) -> dense head on tcgen05 -> PostProcessor
    kernel.  Per-stage times from CUDA events on the launching stream; the head's outputs of a few ROIs are checked
    against the fp32 oracle (stated bf16 tolerance: 3 % of the output range)."""
    import torch

    import oneshotdet_b200 as osd
    from oneshotdet_b200 import ops
    from oracle import oracle as orc

    b, r, c, mlp, h, w = 16, 2000, CHANNELS, 1024, 800, 1344
    g = torch.Generator(device=dev).manual_seed(4242)
    feats = [torch.empty((b, c, -(-h // s), -(-w // s)), device=dev).normal_(generator=g) for s in (8, 16, 32, 64, 128)]
    ctr = torch.rand((b, r, 2), device=dev, generator=g) * torch.tensor([1333.0, 800.0], device=dev)
    wh = torch.rand((b, r, 2), device=dev, generator=g) * 400.0 + 16.0
    rois = torch.cat((ctr - wh / 2, ctr + wh / 2), 2).clamp_(min=0.0)
    rois[..., 2].clamp_(max=1332.0)
    rois[..., 3].clamp_(max=799.0)
    rois = rois.contiguous()
    supp = torch.empty((b, 1, c, 7, 7), device=dev).normal_(generator=g)
    torch.manual_seed(77)
    mods = orc.make_box_head_modules(c, mlp)
    head = osd.BoxHeadDense(c, mlp)
    head.load_state_dict({("predictor." + k if k.startswith(("cls_score", "bbox_pred")) else k): v
                          for k, v in mods.state_dict().items()})
    head = head.to(dev).eval()
    pooler = osd.Pooler((7, 7), [1 / s for s in (8, 16, 32, 64, 128)], 2)
    sizes = [(800, 1333)] * b

    def step():
        rows = pooler.forward_fixed(feats, rois, rows_bf16=True)
        logits, reg = head(rows, supp)
        return rows, logits, reg, ops.box_postprocess(logits, reg, rois, sizes, score_thresh=0.05, nms_thresh=0.5,
                                                      detections_per_img=100)

    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    n = max(3, min(steps, 7))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tp, th, tq = [], [], []
    for _ in range(n):
        ev[0].record()
        rows = pooler.forward_fixed(feats, rois, rows_bf16=True)
        ev[1].record()
        logits, reg = head(rows, supp)
        ev[2].record()
        post = ops.box_postprocess(logits, reg, rois, sizes, score_thresh=0.05, nms_thresh=0.5, detections_per_img=100)
        ev[3].record()
        torch.cuda.synchronize()
        tp.append(ev[0].elapsed_time(ev[1]))
        th.append(ev[1].elapsed_time(ev[2]))
        tq.append(ev[2].elapsed_time(ev[3]))
    # medians: a step that has to go to cudaMalloc for one of its 0.8 GB outputs would otherwise set the mean
    t_pool, t_head, t_post = statistics.median(tp), statistics.median(th), statistics.median(tq)
    # parity spot check of the dense head: 6 ROIs of episode 0 through the fp32 oracle on the fp32 pooled features
    k = 6
    pooled32 = pooler.forward_fixed(feats, rois)[0:1, :k].cpu()
    ol, orr = orc.box_head_dense(pooled32, supp[0:1].cpu(), mods)
    gl, gr = logits[:k].cpu(), reg[:k].cpu()
    err_l = float((gl - ol).abs().max()) / max(float(ol.abs().max()), 1e-12)
    err_r = float((gr - orr).abs().max()) / max(float(orr.abs().max()), 1e-12)
    total = t_pool + t_head + t_post
    flops = 2.0 * b * r * (49 * 4 * c * c + 49 * 2 * c * c + 49 * 9 * c * (c // 2) + 49 * (c // 2) * mlp + mlp * mlp + mlp * 10)
    return {"workload": "second stage at the first stage's output size: 16 episodes x 2000 proposals, 800x1333, C=256, MLP 1024: "
                        "ROI pooler (7x7, sampling 2, 5 levels) -> dense head (concat + compress_dim_conv + feature_aggreg + fc6/fc7 "
                        "+ predictor, tcgen05) -> PostProcessor (softmax, decode, 0.05 / NMS 0.5 / 100)",
            "episodes_per_gpu": b, "rois": b * r, "ms_per_step": total, "episodes_per_s": b / (total * 1e-3), "steps": n,
            "stages": {"pooler_ms": t_pool, "dense_head_ms": t_head, "post_ms": t_post},
            "dense_head_tflops_useful": flops / (t_head * 1e-3) / 1e12,
            "detections_per_episode": post.count.cpu().tolist()[:4],
            "parity": {"checker": "oracle (fp32) on the first 6 ROIs of episode 0", "logits_rel_err": err_l,
                       "regression_rel_err": err_r, "tolerance": 0.03, "ok": bool(err_l <= 0.03 and err_r <= 0.03)}}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from oneshotdet_b200 import ops
    from oneshotdet_b200.distributed import BlockGatherer, DetectionGatherer, PeerBlockGatherer
    from oneshotdet_b200.pipeline import HostStreamer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    warmup = max(args.warmup, 3)
    steps = args.steps
    name = args.workload
    wl = WORKLOADS[name]
    batch = wl["batch"]

    block_mode = args.gather in ("peer", "peer-kernel", "peer-graph", "block")
    overlap = not args.serial
    # steps in flight: the post-processing chains of `depth` consecutive steps run side by side under `depth` matching
    # launches (one chain alone takes longer than one matching launch).  The NCCL block gather has two slots: depth 1.
    depth = 1 if (args.serial or (world > 1 and args.gather in ("block", "packed"))) else max(1, args.depth)
    # --pipeline streams: no per-step graph; three streams that never join between steps (two chains in flight)
    streamed = args.pipeline == "streams" and not args.serial and not args.no_graph and not (world > 1 and args.gather == "packed")
    if streamed:
        depth = 1
    pipe = make_pipe(name, dev, double_buffer=(world > 1 and block_mode), depth=max(2, args.chains) if streamed else depth)
    fill_inputs(pipe, seed=2000 + rank, heads=wl["heads"])
    ep_off = rank * batch
    runner = None if streamed else StepRunner(pipe, depth, overlap=overlap, use_graph=not args.no_graph)

    # N > 1: the step's detections go to every rank.  "peer" (default): the result block of a step (one contiguous
    # buffer of a double-buffered pipeline) is pushed into every rank's receive buffer through CUDA-IPC-mapped peer
    # memory on the copy engines -- no collective kernel shares the SMs with the step.  "peer-kernel": the same push as
    # one small kernel of ours storing through the peer pointers.  "block": one asynchronous NCCL all-gather of the block
    # per step.  "packed": [E, K+1, 6] payloads gathered in groups of --gather-every steps.
    gatherer = None
    K = pipe.post.plan.out_capacity
    if world > 1 and args.gather != "none":
        if args.gather in ("peer", "peer-kernel", "peer-graph"):
            gatherer = PeerBlockGatherer(batch, K, dev, slots=len(pipe.posts), mode="kernel" if args.gather == "peer-kernel" else "copy")
        elif args.gather == "block":
            gatherer = BlockGatherer(batch, K, dev)
        else:
            gatherer = DetectionGatherer(batch, K, dev, ep_off, steps_per_gather=args.gather_every)

    # --gather peer-graph: the push of a step's block is recorded at the end of that step's chain graph (no per-step host
    # work for the exchange at any N).  Measured at N = 2: 213.6 k episodes/s against 218.4 k for the default, which issues
    # the copies from the host on a side stream -- inside the graph they sit on the chain's own stream and delay the next
    # chain of that stream.
    in_graph = streamed and gatherer is not None and args.gather == "peer-graph"
    if streamed:
        hook = (lambda k, r: gatherer.push_on_current_stream(k, r.block)) if in_graph else None
        try:
            runner = StreamRunner(pipe, hook)
        except Exception as exc:  # noqa: BLE001  (capture of the peer copies refused: fall back to per-step submits)
            if not in_graph:
                raise
            print(f"[bench] capturing the peer pushes failed ({exc}); submitting them per step", file=sys.stderr)
            torch.cuda.synchronize()
            in_graph = False
            runner = StreamRunner(pipe)
    launch = runner.launch

    def before(n):
        if in_graph:
            return
        # the exchanges that still read the blocks these steps overwrite have finished.  Full-depth replays cycle through
        # the 2 x depth output sets in submission order (the previous replay's pushes may stay in flight); a remainder
        # step uses the one-step graphs' own rotation, so it waits for everything.
        if gatherer is not None and block_mode:
            if args.gather == "block" or streamed:
                gatherer.acquire(n)    # streamed: output sets rotate in submission order, slots - 1 pushes may stay in flight
            else:
                gatherer.acquire(n, in_flight=(depth if n == depth else 0))

    def after(results):
        if in_graph:
            gatherer.i += len(results)     # bookkeeping only: the pushes are part of the chains' graphs
            return
        if gatherer is not None:       # asynchronous: step i crosses NVLink while step i+1 computes
            for r in results:
                if block_mode:
                    gatherer.submit(r.block)
                else:
                    gatherer.submit(r.boxes, r.scores, r.count)

    res = runner.run(max(warmup, 2 * depth), before, after)
    if gatherer is not None:
        gatherer.finish()
    torch.cuda.synchronize()
    # kernels of this library per step, counted on one eager step (a CUDA-graph replay re-launches the same kernels
    # without passing through the library's launch counter)
    ops.reset_launch_count()
    pipe.run()
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    # isolated stage times (serial, one stream): the roofline of the matching kernel and the post-processing chain
    # (each stage launched back to back n times between two events on the launching stream: a device that idles between
    # single launches -- a host synchronisation after every one -- runs the first microseconds of the next kernel slower)
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    n_iso = max(5, min(args.steps, 20))
    pipe.match()
    pipe.post()
    iso[0].record()
    for _ in range(n_iso):
        pipe.match()
    iso[1].record()
    torch.cuda.synchronize()
    iso[2].record()
    for _ in range(n_iso):
        pipe.post()
    iso[3].record()
    torch.cuda.synchronize()
    iso_match = [iso[0].elapsed_time(iso[1]) / n_iso]
    iso_post = [iso[2].elapsed_time(iso[3]) / n_iso]
    kept = res.kept_before_cut().cpu().tolist()
    counts = res.count.cpu().tolist()

    # ---- timed region: exactly `steps` steps, CUDA events on the launching stream, barrier + sync both sides
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 2)]
    sampler = ClockSampler(local_rank)
    final_gather = None
    if world > 1 and block_mode:
        final_gather = torch.empty((world, gatherer.block_bytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(final_gather.view(-1), res.block)     # NCCL communicator set up before the timing
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ops.reset_launch_count()
    sampler.start()
    t_host0 = time.perf_counter()
    ev[0].record()
    res = runner.run(steps, before, after, mark=lambda done: ev[done].record())
    if gatherer is not None:
        # the last exchanges are inside the timed region: the launching stream waits (on the device) for every push /
        # collective this rank has issued; no host synchronisation and no rendezvous before the end event
        (gatherer.drain if hasattr(gatherer, "drain") else gatherer.finish)()
    ev[steps + 1].record()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    if gatherer is not None:
        gatherer.finish()     # every rank's pushes have landed everywhere (process-group barrier)
        if final_gather is not None:
            # north_star's "final NCCL all-gather of detections": the last step's block, once, through NCCL -- the checker
            # of the pushed copy below (outside the timed region: it is not part of a step)
            dist.all_gather_into_tensor(final_gather.view(-1), res.block)
            torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = launches_per_step * steps
    total_ms = ev[0].elapsed_time(ev[steps + 1])
    marks = [0] + [i for i in range(1, steps + 1) if (i % depth == 0 and i <= steps // depth * depth) or i > steps // depth * depth]
    if streamed:
        marks = [0, steps]     # steps overlap: there is no per-step boundary to mark
    per_step = [ev[a].elapsed_time(ev[b]) / (b - a) for a, b in zip(marks, marks[1:])]
    t = torch.tensor([total_ms, statistics.median(per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max, median_ms_max = float(t[0].item()), float(t[1].item())
    value = batch * n_gpus * steps / (total_ms_max * 1e-3)

    # ---- the exchange delivered every rank's block: pushed copy == NCCL copy == (for this rank) the local block
    exchange = None
    if world > 1 and block_mode:
        slot = (gatherer.i - 1) % (2 if args.gather == "block" else gatherer.slots)
        if args.gather == "block":
            got = gatherer.out[slot].view(world, gatherer.block_bytes)
        else:
            got = gatherer.recv[slot]
        ok_nccl = bool(torch.equal(got, final_gather))
        ok_local = bool(torch.equal(got[rank], res.block))
        flag = torch.tensor([int(ok_nccl and ok_local)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange = {"mode": args.gather, "issued": ("inside the CUDA graph of each step's post-processing chain" if in_graph
                                                    else "per step from the host"),
                    "block_bytes": gatherer.block_bytes, "ranks": world,
                    "equals_nccl_all_gather": ok_nccl, "own_block_round_trip": ok_local, "all_ranks_ok": bool(flag.item())}

    # ---- roofline of the dominant streaming kernel (match_product_bulk_kernel: one launch per step)
    locs = sum(h * w for h, w in pipe.shapes)
    match_bytes = 2 * 4 * CHANNELS * locs * batch                     # fp32 in + out, SURVEY section 8(d)
    match_avg_ms = statistics.mean(iso_match)
    peak, tpeak, peak_src = measured_peaks()
    achieved = match_bytes / (match_avg_ms * 1e-3) / 1e9
    roofline = {"kernel": "match_product_bulk_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                "algorithmic_bytes_per_launch": match_bytes, "avg_launch_ms": match_avg_ms, "peak_source": peak_src,
                "timing": "CUDA events on the launching stream around n back-to-back launches of the kernel running alone, "
                          "right before the timed region; inside the timed region it overlaps the post-processing chain",
                "share_of_serial_step": match_avg_ms / (match_avg_ms + statistics.mean(iso_post))}
    stages = {"match_ms_isolated": match_avg_ms, "post_ms_isolated": statistics.mean(iso_post),
              "serial_ms_per_step": match_avg_ms + statistics.mean(iso_post),
              "overlap": ("matching launches back to back on one stream || the post-processing chains of consecutive steps "
                          "alternating between two more streams, no join between steps (software pipelining across steps)")
              if streamed else (f"match || post-processing on {1 + depth} streams, {depth} step(s) in flight per graph "
                                "replay (software pipelining)") if overlap else "none (one stream)",
              "steps_in_flight": depth, "launch": launch, "post_algorithmic_read_bytes": 6 * 4 * locs * batch,
              "host_ms_per_step": 1e3 * (t_host1 - t_host0) / steps}

    # ---- the 1x1 fusion-conv matching mode (BASELINE configs[1]: "fp32 matching + bf16 1x1 fusion conv"), timed as
    #      its own stage on the same resident features
    fusion = None
    if not args.no_fusion:
        from oneshotdet_b200 import MatchingModule
        from oneshotdet_b200.fusion import PreparedFusion

        torch.manual_seed(1234 + rank)
        mm = MatchingModule("fusion", channels=CHANNELS).to(dev)
        pf = PreparedFusion(pipe.features, pipe.supp, batch, mm.compress_dim_conv, "full")
        for _ in range(3):
            pf()
        torch.cuda.synchronize()
        fsteps = max(3, min(steps, 20))
        f_ms, f_med = time_steps(pf, fsteps, torch)
        flops = 2.0 * locs * batch * (CHANNELS * 2 * CHANNELS + 2 * CHANNELS * CHANNELS)      # executed once (support half folded)
        flops_exec = flops * 1.5                                                                # conv1 runs in both passes
        flops_ref = 2.0 * locs * batch * ((2 * CHANNELS) ** 2 + 2 * CHANNELS * CHANNELS)       # reference form
        hbm_alg = 4.0 * locs * batch * CHANNELS * 2                                             # read x, write out
        hbm_design = 2.0 * locs * batch * CHANNELS * 10                                         # 20C bytes per pixel over the 3 passes
        fusion = {"ms_per_step": f_ms, "ms_per_step_median": f_med, "episodes_per_s": batch / (f_ms * 1e-3), "steps": fsteps,
                  "useful_tflops": flops / (f_ms * 1e-3) / 1e12, "executed_tflops": flops_exec / (f_ms * 1e-3) / 1e12,
                  "reference_form_tflops": flops_ref / (f_ms * 1e-3) / 1e12,
                  "tensor_peak_tflops": tpeak, "tensor_frac_executed": flops_exec / (f_ms * 1e-3) / 1e12 / tpeak,
                  "tensor_peak_source": peak_src, "hbm_algorithmic_bytes": hbm_alg, "hbm_design_bytes": hbm_design,
                  "hbm_gbs_design": hbm_design / (f_ms * 1e-3) / 1e9, "hbm_frac_design": hbm_design / (f_ms * 1e-3) / 1e9 / peak,
                  "what": "compress_dim_conv on P3-P7 for the batch: folded-bias kernel; pass A = conv1 statistics (tcgen05) + "
                          "bf16 copy of x; pass B = conv1 -> GN1 -> LeakyReLU -> conv2 back to back on tcgen05 with the 2C "
                          "intermediate in TMEM; pass C = GroupNorm-2/LeakyReLU in place"}

    # ---- end-to-end through the public API with pinned host buffers: pipelined HostStreamer (H2D of batch i+1 on a copy
    #      stream while batch i replays, D2H on a third stream) and the plain serial run_host, against the measured
    #      host->device ceiling of the same buffers
    e2e = None
    if (rank == 0 or world > 1) and not args.no_e2e:
        host_in = pipe.make_host_inputs(pinned=True)
        for h, d in zip(host_in, pipe.input_tensors()):
            h.copy_(d)
        e2e_steps = max(4, min(steps, 10))
        pipe.run_host(host_in)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = pipe.run_host(host_in)
        torch.cuda.synchronize()
        dt_serial = time.perf_counter() - t0
        ceiling = h2d_ceiling(pipe, host_in, e2e_steps, torch, dist, world, dev)
        pipe2 = make_pipe(name, dev)
        streamer = HostStreamer([pipe, pipe2], use_graph=not args.no_graph)
        for _ in range(2):
            streamer.submit(host_in)
        streamer.drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            streamer.submit(host_in)
        out = streamer.drain()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, dt_serial], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_serial = float(tt[0].item()), float(tt[1].item())
        gbs = pipe.h2d_bytes * e2e_steps / dt / 1e9
        e2e = {"value": batch * n_gpus * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
               "api": "HostStreamer.submit (pinned host inputs -> H2D on a copy stream || graph replay of match + "
                      "post-process || D2H of the detections; one event sync per returned result)",
               "h2d_gbs_per_gpu": gbs, "h2d_ceiling_gbs_per_gpu": ceiling, "frac_of_h2d_ceiling": gbs / ceiling,
               "serial_run_host_value": batch * n_gpus * e2e_steps / dt_serial,
               "check_count0": int(out[2][0])}
        del streamer, pipe2

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload, same inputs as the parity spot check
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = pipe.run()
        torch.cuda.synchronize()
        parity, _ = parity_spot_check(name, pipe, res, with_cpu_time=False)
        ref = CpuReference(name, episodes=1)
        ref.step()
        ts = [sum(ref.step()) for _ in range(4)]
        single = {"value": ref.episodes / statistics.median(ts), "unit": UNIT, "cores": ref.cores, "processes": 1,
                  "sample": ref.sample_text() + f"; median of 4 steps, {statistics.median(ts):.3f} s/episode"}
        cpu = cpu_baseline_dict(single, ref.kind, None if args.no_multi_process else cpu_multi_process_throughput())

    # ---- the other BASELINE configs and the NMS worst cases (N = 1: parity-test cases, reported beside the headline)
    extra = None
    if rank == 0 and world == 1 and not args.no_workloads:
        extra = {}
        del pipe
        torch.cuda.empty_cache()
        for wname in WORKLOADS:
            if wname == name:
                continue
            try:
                extra[wname] = run_extra_workload(wname, dev, steps, rank, use_graph=not args.no_graph, depth=depth,
                                                  streamed=streamed)
            except Exception as exc:  # noqa: BLE001  (report, do not lose the headline line)
                extra[wname] = {"error": repr(exc)}
        try:
            torch.cuda.empty_cache()
            extra["second_stage"] = run_second_stage(dev, steps)
        except Exception as exc:  # noqa: BLE001
            extra["second_stage"] = {"error": repr(exc)}

    if rank == 0:
        gather_text = None
        if world > 1:
            gather_text = {"none": "none (DIAGNOSIS RUN: not a multi-GPU measurement)",
                           "peer": "result block pushed to every rank's receive buffer over peer memory (copy engines), "
                                   "every step, asynchronous (issued from the host on a side stream behind events); one "
                                   "final NCCL all-gather",
                           "peer-graph": "result block pushed to every rank's receive buffer over peer memory (copy "
                                         "engines), every step, the copies recorded in the CUDA graph of the step's "
                                         "post-processing chain; one final NCCL all-gather",
                           "peer-kernel": "result block pushed to every rank by one store kernel over peer memory, every "
                                          "step, asynchronous; one final NCCL all-gather",
                           "block": "result block, NCCL all-gather every step, asynchronous",
                           "packed": f"packed payloads, NCCL all-gather every {args.gather_every} steps, asynchronous"}[args.gather]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": total_ms_max / steps, "ms_per_step_median": median_ms_max, "higher_is_better": True,
                "scaling": "weak", "gather": gather_text, "exchange_check": exchange,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(n_gpus, name, gather_text),
                "roofline": roofline, "stages": stages, "fusion_mode": fusion, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks,
                "check": {"detections_per_episode": counts[:4], "kept_before_cut": kept[:4], "parity": parity},
                "workloads": extra}
        print(json.dumps(line), flush=True)
    if gatherer is not None and hasattr(gatherer, "close"):
        gatherer.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="headline workload of the line (default: BASELINE configs[1]); the others are reported under "
                         "`workloads` of the default N=1 run")
    ap.add_argument("--no-workloads", action="store_true", help="skip the extra workloads of the N=1 run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-process", action="store_true",
                    help="skip the multi-process variant of the CPU baseline (cpu_baseline.multi_process)")
    ap.add_argument("--ref-worker", type=int, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end measurement (diagnosis runs)")
    ap.add_argument("--serial", action="store_true", help="one stream: matching then post-processing, no overlap")
    ap.add_argument("--pipeline", choices=["streams", "graph"], default="streams",
                    help="streams: matching and the post-processing chains of consecutive steps on three self-ordered streams "
                         "(two chains in flight); graph: one CUDA graph per step (match || chain), steps serialised")
    ap.add_argument("--chains", type=int, default=2,
                    help="--pipeline streams: post-processing chains in flight (streams they alternate between)")
    ap.add_argument("--depth", type=int, default=1,
                    help="consecutive steps issued per graph replay, their post-processing chains side by side (1: one step)")
    ap.add_argument("--gather", choices=["peer", "peer-kernel", "peer-graph", "block", "packed", "none"], default="peer",
                    help="N>1: how a step's detections reach every rank (see run_b200_arm); 'none' is a diagnosis run")
    ap.add_argument("--gather-every", type=int, default=10,
                    help="N>1, --gather packed: steps per NCCL all-gather of the detections (1 = every step)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.ref_worker is not None:
        return run_ref_worker(args.ref_worker, max(1, args.steps))
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
