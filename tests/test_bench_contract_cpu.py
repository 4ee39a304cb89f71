"""bench.py contract pieces that run without a GPU: the reference arm prints one JSON line with the agreed keys, and
the B200 arm refuses to run on a CPU-only box instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "matching+NMS episodes/s" and line["unit"] == "episodes/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port", "port+reference-nms") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    # value = all the host cores: one single-threaded worker process per core; the one-process figure sits beside it
    assert line["cpu_processes"] == cb["cores"] == cb["multi_process"]["processes"] >= 1
    assert cb["single_process"]["processes"] == 1 and cb["single_process"]["value"] > 0
    assert abs(line["ms_per_step"] - 1e3 * line["cpu_processes"] / line["value"]) < 1e-6 * line["ms_per_step"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu():
    r = run_bench("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
