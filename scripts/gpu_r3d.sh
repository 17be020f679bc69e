#!/bin/bash
# resident-weight panels for EPI_DGELU2: parity, then the step with and without them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels_r2.py -q -m gpu -k dgelu2 -x 2>&1 | tail -n 12
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "full_size or fixture or full_gradients" 2>&1 | tail -n 5
SKIP_TESTS=1 STEPS=40 VARIANTS="X=1 TULIP_B200_NO_DGELU2_PANEL=1 X=2 TULIP_B200_NO_DGELU2_PANEL=1,X=2" bash scripts/gpu_full.sh
python - <<'PY'
import json
for n in ("X_1", "TULIP_B200_NO_DGELU2_PANEL_1"):
    d = json.loads(open(f"gpurun_out/bench_kitti32_{n}.json").read().strip().splitlines()[-1])
    k = {x["kernel"]: x for x in d["roofline"]["kernels"]}
    print(n, k["gemm_nt<dgelu>"])
PY
