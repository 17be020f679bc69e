#!/bin/bash
# per-stage split of the step at batch 32 / 16 / 8 / 4: does a stage-0/1 block get cheaper per frame when its tensors fit L2?
mkdir -p gpurun_out
for b in 32 16 8 4; do
  timeout 300 python bench.py --batch $b --steps 40 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-sustained > gpurun_out/r3b_b$b.json 2> gpurun_out/r3b_b$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r3b_b$b.json").read().strip().splitlines()[-1])
print("B=$b", d["ms_per_step"], json.dumps(d["step_split"]))
PY
done
