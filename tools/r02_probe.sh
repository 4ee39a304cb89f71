#!/bin/bash
# round-2 probe: bench line + fusion role-wait profile + CTA-pair A/B
python bench.py > gpurun_out/bench_r02_v2.json 2> gpurun_out/bench_r02_v2.err
OSD_FUSION_PROF=1 python tools/fusion_time.py --steps 5 > gpurun_out/fusion_prof.log 2>&1
OSD_FUSION_2CTA=1 python tools/fusion_time.py --steps 10 > gpurun_out/fusion_2cta.log 2>&1
python tools/fusion_time.py --steps 10 > gpurun_out/fusion_1cta.log 2>&1
wc -c gpurun_out/bench_r02_v2.json
