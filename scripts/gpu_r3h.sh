#!/bin/bash
# one-wave pack_weights kernel: parity through the model fixtures, then A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_eval_path.py -q -m gpu -x -k "fixture or full_gradients or droppath or eval or upsample or mc" 2>&1 | tail -n 3
SKIP_TESTS=1 STEPS=40 VARIANTS="X=1 TULIP_B200_LIB=$PWD/build_ab/libtulip_base.so X=2 TULIP_B200_LIB=$PWD/build_ab/libtulip_base.so,X=2" bash scripts/gpu_full.sh
SKIP_TESTS=1 STEPS=20 CONFIG=large8 VARIANTS="X=1 TULIP_B200_LIB=$PWD/build_ab/libtulip_base.so" bash scripts/gpu_full.sh
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_kitti32_X_1.json"))+sorted(glob.glob("gpurun_out/bench_kitti32_TULIP_B200_LIB*base_so.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    k={x["kernel"]:x for x in d["roofline"]["kernels"]}
    print(f[-40:], k["pack_weights"], d["eval_path"])
PY
