#!/bin/bash
F="--no-fusion --no-cpu-baseline --no-workloads --no-e2e --steps 50"
for p in streams graph streams graph; do python bench.py --pipeline $p $F > gpurun_out/bench_pipe_$p.json 2> gpurun_out/bench_pipe_$p.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pipe_$p.json").read().strip().splitlines()[-1])
    print("$p", round(d["value"]), round(d["ms_per_step"],4), round(d["ms_per_step_median"],4), round(d["stages"]["match_ms_isolated"],4), round(d["stages"]["post_ms_isolated"],4), d["roofline"]["frac"], d["check"]["detections_per_episode"][:2])
except Exception as e:
    print("$p FAILED", e); print(open("gpurun_out/bench_pipe_$p.err").read()[-1500:])
PY
done
python tools/timeline.py --steps 1 2>&1 | tail -16
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 1 $F 2>&1 | tail -c 400
