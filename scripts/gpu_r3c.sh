#!/bin/bash
# 2 GPUs: flat all-reduce vs overlapped (three backward phases) with and without an SM reserve for NCCL
mkdir -p gpurun_out
run() {  # name, extra bench args
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 \
    --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained "$@" > gpurun_out/r3c_$name.json 2> gpurun_out/r3c_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r3c_$name.err").read()[-1500:])
PY
}
run flat
run ov0 --allreduce overlap
run ov8 --allreduce overlap --reserve-sms 8
run ov20 --allreduce overlap --reserve-sms 20
