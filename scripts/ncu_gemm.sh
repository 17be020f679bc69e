#!/bin/bash
# ncu --set full captures of single GEMM launches; reports land in gpurun_out/.   usage: ncu_gemm.sh name M N K epi [name M N K epi ...]
mkdir -p gpurun_out
while [ $# -ge 5 ]; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ --launch-skip 2 --launch-count 1 \
    -o gpurun_out/$1 -f python scripts/one_gemm.py $2 $3 $4 $5 > gpurun_out/$1.log 2>&1
  echo "== ncu $1 rc=$?"
  shift 5
done
