#!/bin/bash
# the driver's GPU gate in one process (python -m pytest tests -m gpu), smoke(), then the bench step; logs under gpurun_out/
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 15 > gpurun_out/gputest_full.log
echo "== pytest -m gpu: $(tail -n 1 gpurun_out/gputest_full.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$? $(tail -n 2 gpurun_out/smoke.log)"
fi
B="python bench.py --config ${CONFIG:-kitti32} --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained"
for v in ${VARIANTS:-"X=1"}; do
  name="$(echo $v | tr '= ,/.' '_____')"
  env $(echo $v | tr ',' ' ') timeout 300 $B > gpurun_out/bench_${CONFIG:-kitti32}_$name.json 2> gpurun_out/bench_${CONFIG:-kitti32}_$name.err
  python - "${CONFIG:-kitti32}_$name" <<'P'
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
done
