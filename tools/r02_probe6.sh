#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_box_head.py tests/test_gpu_fcos.py tests/test_gpu_nms.py tests/test_gpu_box_post.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -4
for m in single mc; do echo "== fusion $m"; OSD_FUSION_MODE=$m python tools/fusion_time.py --steps 10 2>&1 | head -7; done
echo "== gram"; OSD_FUSION_GRAM=1 python tools/fusion_time.py --steps 10 2>&1 | head -7
echo "== head"; python tools/box_head_time.py 2>&1 | tail -10
