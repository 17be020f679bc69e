#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels_r2.py -q -m gpu -x -k "test_head_fwd_bwd" 2>&1 | tail -n 30 > gpurun_out/head_test.log
echo "== head kernel test: $(tail -n 1 gpurun_out/head_test.log)"
grep -q "passed" gpurun_out/head_test.log && ! grep -q "failed" gpurun_out/head_test.log || { cat gpurun_out/head_test.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 5 > gpurun_out/head_model.log
echo "== model: $(tail -n 1 gpurun_out/head_model.log)"
SKIP_TESTS=1 STEPS=60 VARIANTS="X=1 TULIP_B200_NO_FUSED_HEAD_BWD=1" bash scripts/gpu_full.sh
