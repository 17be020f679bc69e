#!/bin/bash
# grouped weight-gradient launch: parity tests, then the bench step with the group on / off and a sweep of the item cost
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tn_group.py -q -m gpu --timeout 120 -x 2>&1 | tail -n 30 > gpurun_out/tn_group_test.log
echo "== tn_group: $(tail -n 1 gpurun_out/tn_group_test.log)"
if [ -z "$SKIP_MODEL" ]; then
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -x 2>&1 | tail -n 30 > gpurun_out/tn_group_model.log
echo "== model: $(tail -n 1 gpurun_out/tn_group_model.log)"
fi
B="python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained"
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
    k = {x["kernel"]: x for x in d["roofline"]["kernels"]}
    tn = k.get("gemm_tn", {})
    print(f"== {n}: {d['value']} fps {d['ms_per_step']} ms e2e {d['e2e']['value']} | gemm_tn {tn.get('ms_per_step')} ms {tn.get('launches_per_step')} launches {tn.get('gbs')} GB/s {tn.get('tflops')} TF/s")
except Exception as e:
    print(f"== {n}: failed {e}")
P
}
for v in ${VARIANTS:-"X=1"}; do run "$(echo $v | tr '= ,' '___')" $(echo $v | tr ',' ' '); done
