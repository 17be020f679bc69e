#!/bin/bash
# ncu --set full captures of single GEMM launches (stage-0 shapes of the KITTI B=32 step); reports land in gpurun_out/.
mkdir -p gpurun_out
cap() {  # name M N K epi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ --launch-skip 2 --launch-count 1 \
    -o gpurun_out/$1 -f python scripts/one_gemm.py $2 $3 $4 $5 > gpurun_out/$1.log 2>&1
  echo "== ncu $1 rc=$?"
}
cap qkv_s0 131072 288 96 0
cap fc1_s0 131072 384 96 1
cap fc2_s0 131072 96 384 2
cap wgrad_s0 131072 384 96 9
