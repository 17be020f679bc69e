#!/bin/bash
# Runs the GPU test files one process each (a faulting kernel poisons its CUDA context), logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "traffic" ]; then
for f in test_gpu_index_ops test_gpu_kernels test_gpu_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 -s -rA 2>&1 | tail -n 120 > gpurun_out/$f.log
  echo "== $f: $(tail -n 1 gpurun_out/$f.log)"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$? $(tail -n 2 gpurun_out/smoke.log)"
fi
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -c 3000 gpurun_out/bench.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "== ncu rc=$?"
fi
if [ "$1" == "traffic" ]; then
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; echo "== traffic rc=$?"
fi
