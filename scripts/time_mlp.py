"""Time the fused MLP half-block against the unfused kernel chain at BASELINE cfg2 stage-0 size (T = 131072, C = 96)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
from tests.test_gpu_mlp import make_params, rnd
T, C = 32 * 16 * 256, 96
p = {k: v.cuda() for k, v in make_params(C, 1).items()}
x = rnd(T, C, seed=2, scale=1.5).cuda().to(torch.bfloat16)
w1, w2 = p["mlp.fc1.weight"].to(torch.bfloat16), p["mlp.fc2.weight"].to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def fused(save):
    return ops.mlp_block(x, p["norm2.weight"], p["norm2.bias"], w1, p["mlp.fc1.bias"], w2, p["mlp.fc2.bias"], None, 1, save)

def chain():
    xn, _ = ops.layernorm(x, p["norm2.weight"], p["norm2.bias"])
    h, _ = ops.linear(xn, w1, p["mlp.fc1.bias"], epilogue=ops.EPI_GELU, save_pre=False)
    return ops.linear(h, w2, p["mlp.fc2.bias"], epilogue=ops.EPI_RESID, aux=x)

def timeit(fn, n=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

flops = 16.0 * T * C * C
for name, fn, byt in (("fused (inference)", lambda: fused(False), 4.0 * T * C), ("fused (training by-products)", lambda: fused(True), 14.0 * T * C),
                      ("unfused chain", chain, 26.0 * T * C)):
    m, b = timeit(fn)
    print(f"{name}: median {m:.1f} us (best {b:.1f}) = {flops / m / 1e6:.0f} TFLOP/s, {byt / m / 1e3:.0f} GB/s algorithmic", flush=True)
