#!/bin/bash
F="--no-fusion --no-cpu-baseline --no-workloads --no-e2e --steps 50"
run() { tag=$1; shift; env "$@" python bench.py $F > gpurun_out/bench_np_$tag.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_np_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["ms_per_step"],4), round(d["stages"]["match_ms_isolated"],4), round(d["stages"]["post_ms_isolated"],4))
PY
}
run all A=1
run p2 OSD_NMS_MAX_PASSES=2
run p3 OSD_NMS_MAX_PASSES=3
run all2 A=1
run p2b OSD_NMS_MAX_PASSES=2
