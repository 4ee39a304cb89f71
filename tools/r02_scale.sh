#!/bin/bash
# round-2 scaling probe on an 8-GPU box: the driver's own launch form per N, plus gather-mode / depth A/B at N=8
OUT=gpurun_out
F="--no-cpu-baseline --no-workloads --no-fusion --steps 50 --warmup 5"
run() { # n tag extra...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 $F "$@" > $OUT/scale_r02c_${tag}.json 2> $OUT/scale_r02c_${tag}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $n $F "$@" > $OUT/scale_r02c_${tag}.json 2> $OUT/scale_r02c_${tag}.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/scale_r02c_${tag}.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("$tag", d["n_gpus"], round(d["value"]), round(d["ms_per_step"],4), round(d.get("ms_per_step_median",0),4), "e2e", round(e["value"]) if e else None, e.get("frac_of_h2d_ceiling"), d.get("exchange_check"))
except Exception as ex:
    print("$tag FAILED", ex)
PY
}
run 1 n1 --no-e2e
run 8 n8 --no-e2e
run 2 n2 --no-e2e
run 4 n4 --no-e2e
run 8 n8_none --gather none --no-e2e
run 8 n8_block --gather block --no-e2e
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -3
