#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "cta_pairs and 300" 2>&1 | tail -n 30 > gpurun_out/cg2_first.log
echo "== first pair test: $(tail -n 1 gpurun_out/cg2_first.log)"
grep -q passed gpurun_out/cg2_first.log || { cat gpurun_out/cg2_first.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -x -k "gemm_nt" 2>&1 | tail -n 30 > gpurun_out/cg2_test.log
echo "== gemm_nt tests: $(tail -n 1 gpurun_out/cg2_test.log)"
for m in 0 2; do TULIP_B200_CG2=$m timeout 300 python scripts/time_nt.py 32 96 > gpurun_out/time_nt_cg$m.log 2>&1; tail -n 1 gpurun_out/time_nt_cg$m.log; done
for m in 0 2; do TULIP_B200_CG2=$m timeout 300 python scripts/time_nt.py 8 192 > gpurun_out/time_nt_large_cg$m.log 2>&1; tail -n 1 gpurun_out/time_nt_large_cg$m.log; done
paste -d'|' <(grep "^| [0-9]" gpurun_out/time_nt_cg0.log | cut -d'|' -f2-3,7-9) <(grep "^| [0-9]" gpurun_out/time_nt_cg2.log | cut -d'|' -f7-9)
