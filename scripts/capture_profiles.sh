#!/bin/bash
# Round-end evidence: tests, bench line, ncu launch list + DRAM traffic of one step, ncu --set full of the top kernels.
bash scripts/gpu_check.sh bench
bash scripts/gpu_check.sh traffic > /dev/null 2>&1; echo "== traffic done"
bash scripts/ncu_gemm.sh full_tn_fc1_s0 131072 384 96 9 full_nt_qkv_s0 131072 288 96 0 full_nt_fc1_s0 131072 384 96 1 full_nt_fc1_s3 2048 3072 768 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_bwd --launch-skip 2 --launch-count 1 -o gpurun_out/full_attn_bwd_s0 -f python scripts/one_attn.py 0 > /dev/null 2>&1; echo "== attn bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_fwd --launch-skip 2 --launch-count 1 -o gpurun_out/full_attn_fwd_s0 -f python scripts/one_attn.py 0 > /dev/null 2>&1; echo "== attn fwd rc=$?"
