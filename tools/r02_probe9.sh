#!/bin/bash
F="--no-fusion --no-cpu-baseline --no-workloads --no-e2e --steps 50"
run() { tag=$1; shift; env "$@" python bench.py $F > gpurun_out/bench_knob_$tag.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_knob_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["value"]), round(d["ms_per_step"],4), round(d["stages"]["match_ms_isolated"],4), round(d["roofline"]["frac"],3))
PY
}
run base A=1
run noevict OSD_MATCH_L2_EVICT_FIRST=0
run static OSD_MATCH_SCHED=static
run ctas140 OSD_MATCH_BULK_CTAS=140
run ctas132 OSD_MATCH_BULK_CTAS=132
run base2 A=1
