#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_box_head.py tests/test_gpu_fcos.py tests/test_gpu_pipeline.py tests/test_gpu_pooler.py -x -q 2>&1 | tail -3
python tools/box_head_time.py 2>&1 | tail -12
for c in 592 2368 4736; do python tools/box_head_time.py --chunk $c 2>&1 | head -1; done
for m in cta cluster; do OSD_FCOS_SELECT=$m python bench.py --no-fusion --no-cpu-baseline --no-workloads --steps 30 > gpurun_out/bench_sel_$m.json 2>/dev/null; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_sel_$m.json").read().strip().splitlines()[-1])
print("$m", round(d["value"]), d["ms_per_step"], d["ms_per_step_median"], d["stages"]["match_ms_isolated"], d["stages"]["post_ms_isolated"])
PY
done
OSD_FCOS_SELECT=cta OSD_TIMELINE=1 python tools/timeline.py 2>&1 | tail -12
ncu --set full --clock-control none --import-source on -k regex:'roi_gemm' --launch-skip 12 -c 6 -o gpurun_out/r02_head python tools/box_head_time.py --batch 2 --rois 1184 --steps 1 > gpurun_out/ncu_head.log 2>&1
ncu -i gpurun_out/r02_head.ncu-rep --page raw --csv > gpurun_out/r02_head_raw.csv 2>/dev/null
rm -f gpurun_out/r02_head.ncu-rep
