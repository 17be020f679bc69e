#!/bin/bash
mkdir -p gpurun_out
for s in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_bwd --launch-skip 2 --launch-count 1 \
    -o gpurun_out/attn_bwd_s$s -f python scripts/one_attn.py $s > gpurun_out/attn_bwd_s$s.log 2>&1
  echo "== ncu attn_bwd_s$s rc=$?"
done
