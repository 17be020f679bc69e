#!/bin/bash
# Round-2 evidence: bench lines of the three configs, ncu launch list with DRAM bytes of one step, per-launch table,
# ncu --set full of the kernels that changed this round.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for c in kitti32 durlar16 large8; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err
  echo "== bench $c rc=$? $(head -c 200 gpurun_out/final_bench_$c.json)"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "== ref rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file gpurun_out/final_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-sustained > gpurun_out/final_ncu_traffic.log 2>&1
echo "== traffic rc=$?"
timeout 300 python scripts/launch_table.py 32 > gpurun_out/final_launch_table.md 2> gpurun_out/final_launch_table.err; echo "== table rc=$? $(head -n 1 gpurun_out/final_launch_table.md)"
bash scripts/ncu_step_kernel.sh final_full_head_bwd "mlp_block_fwd_kernel<1>" 1
