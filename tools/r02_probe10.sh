#!/bin/bash
B="python bench.py --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-fusion --no-workloads --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'match_product_bulk' --launch-skip 4 -c 1 -o gpurun_out/match_r02 $B > gpurun_out/ncu_match.log 2>&1
ncu -i gpurun_out/match_r02.ncu-rep --page raw --csv > gpurun_out/match_r02_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'fcos_select|nms_chunk|nms_merge' --launch-skip 9 -c 3 -o gpurun_out/post_r02 $B > gpurun_out/ncu_post.log 2>&1
ncu -i gpurun_out/post_r02.ncu-rep --page raw --csv > gpurun_out/post_r02_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python tools/timeline.py --steps 1 2>&1 | tail -16
