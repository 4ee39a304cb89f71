#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_box_head.py -x -q 2>&1 | tail -5
echo "== fusion gram"; python tools/fusion_time.py --steps 10 2>&1 | head -8
echo "== fusion conv1-stats"; OSD_FUSION_GRAM=0 python tools/fusion_time.py --steps 10 2>&1 | head -1
echo "== head pairs"; python tools/box_head_time.py 2>&1 | tail -10
echo "== head single"; OSD_BOX_HEAD_CLUSTER=1 python tools/box_head_time.py 2>&1 | head -1
python bench.py > gpurun_out/bench_r02_v3.json 2> gpurun_out/bench_r02_v3.err; wc -c gpurun_out/bench_r02_v3.json
