#!/bin/bash
# 2 GPUs, flat all-reduce: NCCL channel count
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 \
    --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained > gpurun_out/r3g_$name.json 2> gpurun_out/r3g_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3g_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r3g_$name.err").read()[-800:])
PY
}
run default X=1
run min32 NCCL_MIN_CTAS=32
run min64 NCCL_MIN_CTAS=64 NCCL_MAX_CTAS=64
run default2 X=2
