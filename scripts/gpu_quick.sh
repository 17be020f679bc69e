#!/bin/bash
# tests (one process per file) + bench + per-launch table; logs under gpurun_out/
mkdir -p gpurun_out
for f in test_gpu_index_ops test_gpu_kernels test_gpu_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 -x 2>&1 | tail -n 40 > gpurun_out/$f.log
  echo "== $f: $(tail -n 1 gpurun_out/$f.log)"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; head -c 420 gpurun_out/bench.json; echo
timeout 300 python scripts/launch_table.py 32 > gpurun_out/launch_table.md 2> gpurun_out/launch_table.err; echo "== table rc=$?"; head -n 1 gpurun_out/launch_table.md
