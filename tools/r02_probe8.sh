#!/bin/bash
F="--no-fusion --no-cpu-baseline --no-workloads --no-e2e --steps 50"
for v in s5 default s7 s8 default s7; do
  if [ $v = default ]; then unset OSD_B200_LIB; else export OSD_B200_LIB=$PWD/oneshotdet_b200/lib/variants/libosd_b200_$v.so; fi
  python bench.py $F > gpurun_out/bench_ring_$v.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ring_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"]), round(d["ms_per_step"],4), round(d["stages"]["match_ms_isolated"],4), round(d["roofline"]["frac"],3))
PY
done
unset OSD_B200_LIB
timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_box_head.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:'match_product_bulk' --launch-skip 4 -c 1 -o gpurun_out/match_r02 python bench.py --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-fusion --no-workloads --no-e2e > gpurun_out/ncu_match.log 2>&1
ncu -i gpurun_out/match_r02.ncu-rep --page raw --csv > gpurun_out/match_r02_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'fcos_select|nms_chunk|nms_merge' --launch-skip 9 -c 3 -o gpurun_out/post_r02 python bench.py --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-fusion --no-workloads --no-e2e > gpurun_out/ncu_post.log 2>&1
ncu -i gpurun_out/post_r02.ncu-rep --page raw --csv > gpurun_out/post_r02_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
