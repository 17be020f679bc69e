#!/bin/bash
# round-2 third session, call A: loss_item test, e2e A/B, free-GELU upper bound
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -k "loss_item or engine_call or deepcopy" -x 2>&1 | tail -n 5
for v in item loss_item; do
  fl=""; [ $v = item ] && fl="--e2e-item"
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained $fl > gpurun_out/r3a_$v.json 2> gpurun_out/r3a_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_$v.json").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"])
PY
done
TULIP_B200_LIB=$PWD/build_ab/libtulip_free.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained > gpurun_out/r3a_free.json 2> gpurun_out/r3a_free.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_free.json").read().strip().splitlines()[-1])
print("free-gelu", d["ms_per_step"], d["value"])
for k in d["roofline"].get("kernels", [])[:14]: print(k)
PY
