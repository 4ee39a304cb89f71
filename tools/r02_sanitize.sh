#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (dense head, fused fusion conv, streamed pipeline, NMS passes)
S="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
run() { name=$1; shift; timeout 420 $S python -m pytest "$@" -x -q -p no:cacheprovider > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$name.log | tail -2 | tr '\n' ' ')"; }
run head tests/test_gpu_box_head.py -k "layers_against_bf16_emulation or pooler_bf16_rows or executed_reference"
run fusion tests/test_gpu_fusion.py -k "full_module and not baseline and not subprocess"
run pipeline tests/test_gpu_pipeline.py
run fcos tests/test_gpu_fcos.py -k "early_exit or stress or ties"
