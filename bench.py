#!/usr/bin/env python
"""bench.py -- matching + FCOS post-processing + NMS throughput (episodes/s) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one pass of the hot path over one batch of 16 synthetic episodes per GPU (BASELINE.json configs[1]:
siamese FCOS R-50-FPN geometry, 800x1333 target padded to 800x1344, C=256, 1 shot, two-stage post-processing
parameters 0 / 6000 / 0.8 / 2000): the product matching of P3-P7 (one launch), then score -> per-level top-k ->
decode/clip -> batched NMS -> post-NMS top-n.  The FCOS head between the two stages is outside the path: head
outputs are synthetic and resident.  With N > 1 every rank owns 16 episodes (weak scaling) and
every step's detections (the post-processing stage's result block: boxes, scores, indices, counts) go to all ranks
with one asynchronous NCCL all-gather that overlaps the next step (--gather packed: [16, 2001, 6] payloads in groups).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric through
EpisodePipeline.run_host with pinned HOST buffers (H2D of all inputs + D2H of the detections inside the timed
region), `roofline` describes the dominant streaming kernel, `cpu_baseline` the reference's CPU path timed on
this box's host cores on a bounded sample.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "matching+NMS episodes/s"
UNIT = "episodes/s"
BATCH = 16
HEIGHT, WIDTH = 800, 1344          # 800x1333 zero-padded to a multiple of 32 (structures/image_list.py:56-63)
IMAGE_SIZE = (800, 1333)
CHANNELS, SHOTS = 256, 1
PARAMS = dict(pre_nms_thresh=0.0, pre_nms_top_n=6000, nms_thresh=0.8, fpn_post_nms_top_n=2000, min_size=0.0)
WORKLOAD = ("configs[1]: siamese FCOS R-50-FPN geometry, 16 episodes/GPU, 800x1333 (padded 800x1344), C=256, "
            "1 shot, fp32 product matching + score/top-k(6000)/decode/NMS(0.8)/top-2000")


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "episodes_per_gpu": BATCH, "global_episodes": BATCH * n_gpus,
            "levels": "P3-P7 100x168,50x84,25x42,13x21,7x11", "match_mode": "product", "match_dtype": "f32",
            "post_params": PARAMS, "parallelism": f"episode-dp{n_gpus}",
            "streams": "2 (matching || post-processing), see stages.overlap",
            "cache": "inputs larger than L2 (367 MB features in + 367 MB out per step vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------------
# synthetic inputs: features ~ N(0,1); cls ~ N(-4, 2^2); ctr ~ N(0,1); reg = exp(N(log 4s, .5^2))
# ------------------------------------------------------------------------------------------------------
def fill_inputs(pipe, seed):
    import math

    import torch

    g = torch.Generator(device=pipe.device).manual_seed(seed)
    for t in pipe.features + pipe.supp:
        t.normal_(generator=g)
    for t in pipe.cls:
        t.normal_(mean=-4.0, std=2.0, generator=g)
    for t in pipe.ctr:
        t.normal_(generator=g)
    for t, s in zip(pipe.reg, pipe.strides):
        t.normal_(mean=math.log(4.0 * s), std=0.5, generator=g).exp_()


# ------------------------------------------------------------------------------------------------------
# the reference's CPU path on a bounded sample: torch.mul matching (generalized_rcnn.py:306-311) + the
# post-processing restated with the same ATen ops (oracle) + the reference's own nms_cpu when it was compiled
# ------------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, episodes=1, seed=7, threads=None):
        import torch

        from oracle import build_ref
        from oracle import oracle as orc

        self.orc, self.torch = orc, torch
        torch.set_num_threads(threads or os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        ref = None
        try:
            ref = build_ref.load_ref()
        except Exception:  # noqa: BLE001
            ref = None
        self.kind = "reference" if ref is not None else "port"
        if ref is not None:
            self.nms_fn = lambda b, s, thr: ref.nms(torch.from_numpy(b), torch.from_numpy(s), float(thr)).numpy()
        else:
            self.nms_fn = None
        self.episodes = episodes
        self.feats, self.supp = orc.synth_features(episodes, SHOTS, CHANNELS, HEIGHT, WIDTH, seed)
        self.cls, self.reg, self.ctr = orc.synth_head_outputs(episodes, HEIGHT, WIDTH, seed, distinct=False)
        self.params = orc.PostParams(**PARAMS)
        self.sizes = [IMAGE_SIZE] * episodes

    def step(self):
        orc = self.orc
        t0 = time.perf_counter()
        out = orc.match_product(self.feats, self.supp, self.episodes)
        t1 = time.perf_counter()
        res = orc.fcos_postprocess(self.cls, self.reg, self.ctr, orc.FPN_STRIDES, self.sizes, self.params,
                                   nms_fn=self.nms_fn)
        t2 = time.perf_counter()
        assert len(out) == 5 and len(res) == self.episodes
        return t1 - t0, t2 - t1

    def sample_text(self):
        nms = ("the reference's own nms_cpu (csrc/cpu/nms_cpu.cpp compiled unmodified, oracle/_ref)"
               if self.kind == "reference" else "the C port of nms_cpu (oracle/nms_oracle.c)")
        return (f"{self.episodes} episode(s) of the same workload per step: torch.mul matching on {self.cores} threads + "
                f"FCOS post-processing restated with the reference's ATen ops + {nms}; nms_cpu is single-threaded by "
                f"construction (no OpenMP)")


def cpu_multi_process_throughput(processes=None, timed_steps=2):
    """The same CPU path run as `processes` independent single-threaded workers (one episode stream each) -- what the host
    cores deliver when the reference's per-image loop is data-parallelised over processes, since its nms_cpu cannot use a
    second thread.  Every worker is this script with --ref-worker; the aggregate rate is the sum of the workers' rates
    (they overlap for the whole timed part: all start by importing torch and running one warm-up step)."""
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
    return {"value": rate, "unit": UNIT, "processes": ok, "threads_per_process": 1,
            "what": "independent single-threaded workers, one episode per step each; sum of the workers' rates"}


def run_ref_worker(seed, steps):
    ref = CpuReference(episodes=1, seed=seed, threads=1)
    ref.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step()
    print(steps * ref.episodes / (time.perf_counter() - t0), flush=True)
    return 0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(episodes=1)
    for _ in range(max(args.warmup, 1)):
        ref.step()
    times = []
    budget_s = 150.0   # keep the whole run within a few minutes whatever K is; `steps` reports what was timed
    for _ in range(args.steps):
        a, b = ref.step()
        times.append(a + b)
        if sum(times) > budget_s:
            break
    args.steps = len(times)
    total = sum(times)
    value = ref.episodes * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                             "sample": ref.sample_text(),
                             "multi_process": None if args.no_multi_process else cpu_multi_process_throughput()},
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


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram__bytes_read+write of the match kernel per launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "match_kernel_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


# ------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from oneshotdet_b200 import ops
    from oneshotdet_b200.distributed import BlockGatherer, DetectionGatherer
    from oneshotdet_b200.pipeline import EpisodePipeline, PostParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL is left at its defaults: fewer channels / the LL protocol help the 643 KB gather at 2 GPUs (0.174 vs
        # 0.182 ms per step) but starve it at 8 (0.26 - 0.46 ms), see DESIGN section 7.
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    warmup = max(args.warmup, 3)
    steps = args.steps

    pipe = EpisodePipeline(BATCH, HEIGHT, WIDTH, [IMAGE_SIZE] * BATCH, channels=CHANNELS, shots=SHOTS,
                           params=PostParams(**PARAMS), match_mode="product", device=dev,
                           double_buffer=(world > 1 and args.gather == "block"))
    fill_inputs(pipe, seed=2000 + rank)
    ep_off = rank * BATCH

    overlap = not args.serial
    launch = "eager"
    step_fn = pipe.run_overlapped if overlap else pipe.run
    if not args.no_graph:
        try:
            step_fn = pipe.capture(overlapped=overlap)
            launch = "cuda-graph"
        except Exception as exc:  # noqa: BLE001  (launch mechanism only; the kernels are the same either way)
            print(f"[bench] CUDA-graph capture failed ({exc}); launching eagerly", file=sys.stderr)
            torch.cuda.synchronize()

    # N > 1: the step's detections go to every rank with one asynchronous NCCL all-gather.  "block" (default): the
    # post-processing outputs of a step are one contiguous result block in a double-buffered pipeline and that block is
    # gathered as is, every step, without packing kernels.  "packed": [E, K+1, 6] payloads packed by copy kernels and
    # gathered in groups of --gather-every steps (the reference gathers once, after the whole dataset).
    gatherer = None
    if world > 1 and args.gather != "none":
        if args.gather == "block":
            gatherer = BlockGatherer(BATCH, pipe.post.plan.out_capacity, dev)
        else:
            gatherer = DetectionGatherer(BATCH, pipe.post.plan.out_capacity, dev, ep_off, steps_per_gather=args.gather_every)

    def step():
        res = step_fn()
        if gatherer is not None:   # asynchronous: NCCL moves step i over NVLink while step i+1 computes
            if args.gather == "block":
                gatherer.submit(res.block)
            else:
                gatherer.submit(res.boxes, res.scores, res.count)
        return res

    for _ in range(warmup):
        res = step()
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
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    iso_match, iso_post = [], []
    for _ in range(max(3, min(args.steps, 10))):
        iso[0].record()
        pipe.match()
        iso[1].record()
        pipe.post()
        iso[2].record()
        torch.cuda.synchronize()
        iso_match.append(iso[0].elapsed_time(iso[1]))
        iso_post.append(iso[1].elapsed_time(iso[2]))
    kept = res.kept_before_cut().cpu().tolist()
    counts = res.count.cpu().tolist()

    # ---- timed region: exactly `steps` steps, CUDA events on the launching stream, barrier + sync both sides
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ops.reset_launch_count()
    sampler.start()
    t_host0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
    if gatherer is not None:
        gatherer.finish()     # the last gathers are inside the timed region
        ev[-1][1].record()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = launches_per_step * steps
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    match_ms, post_ms = iso_match, iso_post
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = BATCH * n_gpus * steps / (total_ms_max * 1e-3)

    # ---- roofline of the dominant streaming kernel (match_product_bulk_kernel: one launch per step)
    locs = sum(h * w for h, w in pipe.shapes)
    match_bytes = 2 * 4 * CHANNELS * locs * BATCH                     # fp32 in + out, SURVEY section 8(d)
    match_avg_ms = statistics.mean(match_ms)
    peak, peak_src = measured_peak_hbm()
    achieved = match_bytes / (match_avg_ms * 1e-3) / 1e9
    roofline = {"kernel": "match_product_bulk_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                "algorithmic_bytes_per_launch": match_bytes, "avg_launch_ms": match_avg_ms, "peak_source": peak_src,
                "timing": "CUDA events around the kernel on its launching stream, kernel running alone (serial loop "
                          "right before the timed region); inside the timed region it overlaps the post-processing chain",
                "share_of_serial_step": match_avg_ms / (match_avg_ms + statistics.mean(post_ms))}
    post_read = 6 * 4 * locs * BATCH
    stages = {"match_ms_isolated": match_avg_ms, "post_ms_isolated": statistics.mean(post_ms),
              "serial_ms_per_step": match_avg_ms + statistics.mean(post_ms),
              "overlap": "match || post-processing on two streams (software pipelining)" if overlap else "none (one stream)",
              "launch": launch,
              "post_algorithmic_read_bytes": post_read,
              "host_ms_per_step": 1e3 * (t_host1 - t_host0) / steps}

    # ---- the 1x1 fusion-conv matching mode (BASELINE configs[1]: "fp32 matching + bf16 1x1 fusion conv"), timed as
    #      its own stage on the same resident features: bias fold + conv1 (tcgen05) + conv2 (tcgen05) + GN/LeakyReLU
    fusion = None
    if not args.no_fusion:
        from oneshotdet_b200 import MatchingModule
        from oneshotdet_b200.fusion import PreparedFusion

        torch.manual_seed(1234 + rank)
        mm = MatchingModule("fusion", channels=CHANNELS).to(dev)
        pf = PreparedFusion(pipe.features, pipe.supp, BATCH, mm.compress_dim_conv, "full")
        for _ in range(3):
            pf()
        torch.cuda.synchronize()
        fsteps = max(3, min(steps, 20))
        fe = [torch.cuda.Event(enable_timing=True) for _ in range(fsteps + 1)]
        fe[0].record()
        for i in range(fsteps):
            pf()
            fe[i + 1].record()
        torch.cuda.synchronize()
        f_ms = fe[0].elapsed_time(fe[-1]) / fsteps
        flops = 2.0 * locs * BATCH * (CHANNELS * 2 * CHANNELS + 2 * CHANNELS * CHANNELS)      # executed (support half folded)
        flops_ref = 2.0 * locs * BATCH * ((2 * CHANNELS) ** 2 + 2 * CHANNELS * CHANNELS)       # reference form
        hbm = 4.0 * locs * BATCH * CHANNELS * (1 + 2 + 2 + 1 + 2)                               # x, y1 w+r, y2 w, GN pass r+w
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                tpeak = float(json.load(f)["bf16_tflops_sustained"])
            tsrc = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:  # noqa: BLE001
            tpeak, tsrc = 1400.0, "fallback (B200_PROFILING.md sustained)"
        fusion = {"ms_per_step": f_ms, "episodes_per_s": BATCH / (f_ms * 1e-3), "steps": fsteps,
                  "executed_tflops": flops / (f_ms * 1e-3) / 1e12, "reference_form_tflops": flops_ref / (f_ms * 1e-3) / 1e12,
                  "tensor_peak_tflops": tpeak, "tensor_frac_executed": flops / (f_ms * 1e-3) / 1e12 / tpeak,
                  "tensor_peak_source": tsrc, "hbm_algorithmic_bytes": hbm,
                  "hbm_gbs": hbm / (f_ms * 1e-3) / 1e9, "hbm_frac": hbm / (f_ms * 1e-3) / 1e9 / peak,
                  "what": "compress_dim_conv on P3-P7 for 16 episodes: folded-bias kernel + 2 tcgen05 GEMM launches "
                          "(bf16 operands, fp32 accumulate, fp32 intermediates) + GroupNorm/LeakyReLU pass"}

    # ---- end-to-end through the public API with pinned host buffers
    e2e = None
    if rank == 0 or world > 1:
        host_in = pipe.make_host_inputs(pinned=True)
        for h, d in zip(host_in, pipe.input_tensors()):
            h.copy_(d)
        e2e_steps = max(3, min(steps, 10))
        pipe.run_host(host_in)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = pipe.run_host(host_in)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": BATCH * n_gpus * e2e_steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps,
               "api": "EpisodePipeline.run_host (pinned host inputs -> H2D -> match + post-process -> D2H detections)",
               "check_count0": int(out[2][0])}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(episodes=1)
        ref.step()
        ts = [sum(ref.step()) for _ in range(4)]
        cpu = {"value": ref.episodes / statistics.median(ts), "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
               "sample": ref.sample_text() + f"; median of 4 steps, {statistics.median(ts):.3f} s/episode",
               "multi_process": None if args.no_multi_process else cpu_multi_process_throughput()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": total_ms_max / steps, "higher_is_better": True, "scaling": "weak",
                "gather": (("none (DIAGNOSIS RUN: not a multi-GPU measurement)" if args.gather == "none" else
                            "result block, every step, async" if args.gather == "block" else
                            f"packed payloads, every {args.gather_every} steps, async") if world > 1 else None),
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(n_gpus),
                "roofline": roofline, "stages": stages, "fusion_mode": fusion, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks,
                "check": {"detections_per_episode": counts[:4], "kept_before_cut": kept[:4]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-process", action="store_true",
                    help="skip the multi-process variant of the CPU baseline (cpu_baseline.multi_process)")
    ap.add_argument("--ref-worker", type=int, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one stream: matching then post-processing, no overlap")
    ap.add_argument("--gather", choices=["block", "packed", "none"], default="block",
                    help="N>1: gather the step's result block as is (default), pack [E,K+1,6] payloads, or (diagnosis "
                         "only: not a valid multi-GPU measurement) no gather at all")
    ap.add_argument("--gather-every", type=int, default=10,
                    help="N>1: steps per NCCL all-gather of the detections (1 = every step)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.ref_worker is not None:
        return run_ref_worker(args.ref_worker, max(1, args.steps))
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
