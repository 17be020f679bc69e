"""Time the fused W-MSA half-block against the unfused kernel chain at BASELINE cfg2 stage-0 size (B=32, 16x256 tokens, C=96)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
from tests.test_gpu_wmsa import make_block_params, rnd

B, H, W, C, heads = int(os.environ.get("B", 32)), 16, 256, 96, 3
p = {k: v.cuda() for k, v in make_block_params(C, heads, 1).items()}
x = rnd(B * H * W, C, seed=2, scale=1.5).cuda().to(torch.bfloat16)
wq, wp = p["attn.qkv.weight"].to(torch.bfloat16), p["attn.proj.weight"].to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def fused(shift):
    return ops.wmsa_block(x, p["norm1.weight"], p["norm1.bias"], wq, p["attn.qkv.bias"], wp, p["attn.proj.bias"],
                          p["attn.relative_position_bias_table"], B, H, W, heads, (2, 8), (1, 4) if shift else (0, 0), shift)

def chain(shift):
    xn, _ = ops.layernorm(x, p["norm1.weight"], p["norm1.bias"])
    qkv = ops.linear(xn, wq, p["attn.qkv.bias"])
    o = ops.window_attention(qkv, p["attn.relative_position_bias_table"], B, H, W, heads, (2, 8), (1, 4) if shift else (0, 0), shift)
    return ops.linear(o, wp, p["attn.proj.bias"], epilogue=ops.EPI_RESID, aux=x)

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

T = B * H * W
flops = 8.0 * T * C * C + 64.0 * T * C
for shift in (False, True):
    mf, bf_ = timeit(lambda: fused(shift))
    mc, bc = timeit(lambda: chain(shift))
    print(f"shift={shift}: fused median {mf:.1f} us (best {bf_:.1f})  = {flops / mf / 1e6:.0f} TFLOP/s, {4.0 * T * C / mf / 1e3:.0f} GB/s algorithmic;"
          f"  unfused chain median {mc:.1f} us (best {bc:.1f})", flush=True)

# phase trace of CTA 0 (clock64 stamps; bring-up hook)
import ctypes as C
from tulip_b200._lib import load_library
lib = load_library()
tr = torch.zeros(4 * 8 * 8, dtype=torch.int64, device="cuda")
lib.tulip_debug_wmsa_trace.argtypes = [C.c_void_p]
lib.tulip_debug_wmsa_trace(tr.data_ptr())
fused(True); torch.cuda.synchronize()
lib.tulip_debug_wmsa_trace(None)
t = tr.cpu().view(4, 8, 8)
t0 = int(t[t > 0].min())
names = {0: ["mma: loop top", "a_full ok", "qkv_empty ok", "o_full ok", "proj issued"],
         1: ["ln: -", "-", "-", "norm start", "raw landed", "a_empty ok", "norm end"],
         3: ["epi(it): start", "p_full ok", "tmem loaded", "chunks done", "fenced+arrived", "S done (compute)", "softmax done"],
         2: ["at: top", "qkv_full ok", "frags loaded", "computed", "o_empty ok", "O stored", "epi(it-1) done"]}
print('entry -> setup done -> t0 -> exit (cycles rel. t0):', [int(v) - t0 for v in t[0, 7][:3]])
for role in range(4):
    for it in range(7):
        row = [(int(v) - t0) if v > 0 else -1 for v in t[role, it][:len(names[role])]]
        print(f"role {role} tile {it}: " + "  ".join(f"{n}={v}" for n, v in zip(names[role], row)))
