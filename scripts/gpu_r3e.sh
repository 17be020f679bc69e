#!/bin/bash
# bias prefetch for EPI_DGELU2 + multiply-high tile index: parity, then A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_kernels_r2.py tests/test_gpu_mlp.py -q -m gpu -x 2>&1 | tail -n 4
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "full_size or fixture" 2>&1 | tail -n 3
SKIP_TESTS=1 STEPS=40 VARIANTS="X=1 TULIP_B200_LIB=$PWD/build_ab/libtulip_base.so X=2 TULIP_B200_LIB=$PWD/build_ab/libtulip_base.so,X=2" bash scripts/gpu_full.sh
