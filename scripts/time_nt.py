"""Forward / dX GEMMs of the Swin blocks, stage by stage, replayed from a CUDA graph (no host launch cost, no per-launch
events).  Compare schedules by running it under TULIP_B200_CG2=0 / 1.  usage: time_nt.py [batch] [embed_dim]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from tulip_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
E0 = int(sys.argv[2]) if len(sys.argv) > 2 else 96
REP = 40
lib = ops.load_library()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(REP):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / REP


print(f"CG2={os.environ.get('TULIP_B200_CG2', 'default')}  batch {B} embed {E0}")
print("| stage | gemm | M | N | K | schedule | us | TFLOP/s |")
print("|---|---|---:|---:|---:|---|---:|---:|")
tot = 0.0
for s in range(4):
    T, Cc = (B * 4096) >> (2 * s), E0 << s
    for name, N, K, epi in (("qkv", 3 * Cc, Cc, 0), ("proj", Cc, Cc, 2), ("fc1", 4 * Cc, Cc, 1), ("fc2", Cc, 4 * Cc, 2),
                            ("dX qkv", Cc, 3 * Cc, 0), ("dX fc1", Cc, 4 * Cc, 0), ("dX proj", Cc, Cc, 0)):
        x = torch.randn(T, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
        b = torch.randn(N, device="cuda")
        aux = torch.randn(T, N, device="cuda").bfloat16()
        if epi == 0:
            fn = lambda: ops.linear(x, w, b, impl=2)
        elif epi == 1:
            fn = lambda: ops.linear(x, w, b, epilogue=1, impl=2, save_pre=False)
        else:
            fn = lambda: ops.linear(x, w, b, epilogue=2, aux=aux, impl=2)
        out = (C.c_int * 10)()
        lib.tulip_gemm_nt_plan(T, N, K, epi, 0, out)
        t = timed(fn)
        tot += t
        print(f"| {s} | {name} | {T} | {N} | {K} | bn{out[0]} {('tile', 'panel', 'pairs')[out[1]]} grid {out[8]} | {t:.1f} | {2.0 * T * N * K / t / 1e6:.0f} |")
print(f"sum {tot:.1f} us")
