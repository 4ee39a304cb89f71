#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_box_head.py -x -q 2>&1 | tail -2
python tools/box_head_time.py 2>&1 | tail -10
python tools/box_head_time.py --chunk 4736 2>&1 | head -1
for d in 1 2 3; do python bench.py --depth $d --no-fusion --no-cpu-baseline --no-workloads --steps 50 > gpurun_out/bench_depth$d.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_depth$d.json").read().strip().splitlines()[-1])
print("depth $d", round(d["value"]), round(d["ms_per_step"],4), round(d["ms_per_step_median"],4), round(d["stages"]["match_ms_isolated"],4), round(d["stages"]["post_ms_isolated"],4))
PY
done
python __graft_entry__.py --smoke 2>&1 | tail -2
