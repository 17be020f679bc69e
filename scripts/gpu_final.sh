#!/bin/bash
# the driver's round-end sequence on one box: pytest -m gpu, smoke(), the reference arm, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 15 > gpurun_out/gputest_full.log
echo "== pytest -m gpu: $(tail -n 1 gpurun_out/gputest_full.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$? $(tail -n 2 gpurun_out/smoke.log)"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err ) 2>&1 | grep real
echo "== reference arm rc=$? $(head -c 300 gpurun_out/final_bench_ref.json)"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real
echo "== bench rc=$? $(head -c 600 gpurun_out/final_bench.json)"
for c in durlar16 large8; do
  timeout 900 python bench.py --config $c --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err
  echo "== bench $c rc=$? $(head -c 260 gpurun_out/final_bench_$c.json)"
done
