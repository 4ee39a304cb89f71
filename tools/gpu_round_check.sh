#!/bin/bash
# One-GPU round check: the driver's own sequence (GPU tests, smoke, both bench arms) plus the ncu launch list of the
# bench command and `--set full` captures of the path's kernels.  Outputs land in gpurun_out/ (copy what should be judged
# into profiles/).  Usage: bash tools/gpu_round_check.sh [tag]
set -x
TAG=${1:-r02_final}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>$O/bench_$TAG.err | tail -1 > $O/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference_arm_$TAG.json
# launch list of the bench command (serialised, cold-cache per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-workloads --no-e2e > $O/ncu_launch.log 2>&1
# full captures: the streaming / post-processing kernels of one step
ncu --set full --clock-control none --import-source on -k regex:'match_product_bulk|nms_mask|fcos_select|nms_sweep|nms_merge|nms_chunk' \
    --launch-skip 40 -c 6 -o $O/full_$TAG python bench.py --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-fusion --no-workloads --no-e2e > $O/ncu_full.log 2>&1
ncu -i $O/full_$TAG.ncu-rep --page raw --csv > $O/full_${TAG}_raw.csv 2>/dev/null
# the NMS kernels on the full problem (early_exit=False: all 67.3 M pairs per episode)
ncu --set full --clock-control none --import-source on -k regex:'nms_mask|nms_sweep' --launch-skip 8 -c 2 \
    -o $O/nmsfull_$TAG python bench.py --workload nms_noexit --steps 2 --warmup 3 --pipeline graph --serial --no-graph --no-cpu-baseline --no-fusion --no-workloads --no-e2e > $O/ncu_nmsfull.log 2>&1
ncu -i $O/nmsfull_$TAG.ncu-rep --page raw --csv > $O/nmsfull_${TAG}_raw.csv 2>/dev/null
# the fused fusion-conv kernels
ncu --set full --clock-control none --import-source on -k regex:'fusion_stats1|fusion_b2b|fusion_gn_lrelu' --launch-skip 3 -c 3 \
    -o $O/fusion_$TAG python tools/fusion_time.py --steps 2 > $O/ncu_fusion.log 2>&1
ncu -i $O/fusion_$TAG.ncu-rep --page raw --csv > $O/fusion_${TAG}_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
cut -c1-300 $O/bench_$TAG.json
# the dense head's GEMM launches of one chunk
ncu --set full --clock-control none --import-source on -k regex:'roi_gemm' --launch-skip 12 -c 6 -o $O/head_$TAG python tools/box_head_time.py --batch 2 --rois 1184 --steps 1 > $O/ncu_head.log 2>&1
ncu -i $O/head_$TAG.ncu-rep --page raw --csv > $O/head_${TAG}_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
