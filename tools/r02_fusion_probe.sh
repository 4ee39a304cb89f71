#!/bin/bash
# fusion launch forms: timing + parity
for m in single mc pair; do
  echo "== $m"; OSD_FUSION_MODE=$m timeout 300 python tools/fusion_time.py --steps 10 2>&1 | grep -v "fusion prof" | head -12
done
timeout 900 python -m pytest tests/test_gpu_fusion.py -x -q 2>&1 | tail -5
