set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01_v8.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_reference_arm_v8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_launches_v8.csv python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'match_product_bulk|conv1x1_tc|nms_mask|fcos_select|nms_sweep|nms_merge|nms_chunk' --launch-skip 40 -c 12 -o gpurun_out/r01_full_v8 python bench.py --steps 2 --warmup 3 --serial --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r01_full_v8.ncu-rep --page raw --csv > gpurun_out/r01_full_v8_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
cat gpurun_out/bench_r01_v8.json | cut -c1-400
