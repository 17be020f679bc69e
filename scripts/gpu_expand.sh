#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -s -k "expanding or base_kitti" 2>&1 | tail -n 30 > gpurun_out/expand_test.log
echo "== expanding: $(tail -n 1 gpurun_out/expand_test.log)"; grep "^\[model" gpurun_out/expand_test.log
