#!/bin/bash
# One-GPU round check: the driver's own sequence (GPU tests, smoke, both bench arms) plus the ncu launch list and one
# `--set full` capture of the path's kernels.  Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -x
TAG=${1:-v10}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_reference_arm_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_launches_$TAG.csv python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'match_product_bulk|nms_mask|fcos_select|nms_sweep|nms_merge|nms_chunk' --launch-skip 40 -c 6 -o gpurun_out/r01_full_$TAG python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline --no-fusion > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r01_full_$TAG.ncu-rep --page raw --csv > gpurun_out/r01_full_${TAG}_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'conv1x1_tc|fusion_gn_lrelu' --launch-skip 2 -c 3 -o gpurun_out/r01_fusion_$TAG python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline > gpurun_out/ncu_fusion.log 2>&1
ncu -i gpurun_out/r01_fusion_$TAG.ncu-rep --page raw --csv > gpurun_out/r01_fusion_${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cut -c1-200 gpurun_out/bench_r01_$TAG.json
