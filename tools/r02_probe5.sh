#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_box_head.py tests/test_gpu_fcos.py tests/test_gpu_nms.py tests/test_gpu_pipeline.py tests/test_gpu_box_post.py -x -q 2>&1 | tail -4
echo "== fusion gram"; python tools/fusion_time.py --steps 10 2>&1 | head -8
echo "== head"; python tools/box_head_time.py 2>&1 | tail -10
python bench.py > gpurun_out/bench_r02_v4.json 2> gpurun_out/bench_r02_v4.err; wc -c gpurun_out/bench_r02_v4.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_v4.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["roofline"]["frac"], d["stages"]["match_ms_isolated"], d["stages"]["post_ms_isolated"], d["fusion_mode"]["ms_per_step"])
for k,v in d["workloads"].items(): print(k, round(v["ms_per_step"],4), v["stages"], v["parity"]["ok"])
PY
