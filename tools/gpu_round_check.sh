#!/bin/bash
# One-GPU round check: the driver's own sequence (GPU tests, smoke, both bench arms) plus the ncu launch list.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_v9.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_reference_arm_v9.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_launches_v9.csv python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
cut -c1-300 gpurun_out/bench_r01_v9.json
