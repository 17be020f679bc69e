"""Bring-up check of the tcgen05 GEMM against torch (fp32) on a few shapes; prints error patterns."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops

torch.manual_seed(0)
def bf(t): return t.to(torch.bfloat16)
shapes = [(128, 96, 64), (128, 96, 96), (256, 192, 128), (1000, 288, 96), (131072, 288, 96), (131072, 384, 96), (131072, 96, 384), (2048, 2304, 768), (64, 96, 96)]
for (M, N, K) in shapes:
    x = bf(torch.randn(M, K, device="cuda")); w = bf(torch.randn(N, K, device="cuda") * K ** -0.5); b = torch.randn(N, device="cuda")
    want = (x.float() @ w.float().t() + b)
    for impl in (2,):
        got = ops.linear(x, w, b, impl=impl).float()
        torch.cuda.synchronize()
        err = (got - want).abs()
        rel = ((got - want).norm() / want.norm()).item()
        print(f"M={M} N={N} K={K} impl={impl}: rel {rel:.3e} max {err.max().item():.3e}", flush=True)
        if rel > 1e-2:
            bad = (err > 0.1).nonzero()
            print("   bad count", bad.shape[0], "first", bad[:5].tolist(), "rows bad:", sorted(set((bad[:, 0] % 128).tolist()))[:20],
                  "cols bad:", sorted(set(bad[:, 1].tolist()))[:20])
# timing
for (M, N, K) in [(131072, 288, 96), (131072, 384, 96), (131072, 96, 384), (32768, 576, 192), (2048, 2304, 768), (2048, 768, 3072)]:
    x = bf(torch.randn(M, K, device="cuda")); w = bf(torch.randn(N, K, device="cuda") * K ** -0.5); b = torch.randn(N, device="cuda")
    for impl, name in ((1, "mma"), (2, "tc05")):
        for _ in range(3): ops.linear(x, w, b, impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.linear(x, w, b, impl=impl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"M={M} N={N} K={K} {name}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s  {(M*K+N*K+M*N)*2/ms/1e6:.0f} GB/s", flush=True)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3): torch.nn.functional.linear(x, w)
    t0.record()
    for _ in range(20): torch.nn.functional.linear(x, w)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 20
    print(f"M={M} N={N} K={K} cublas: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
print("---- TN (weight gradient) ----")
for (M, N, K) in [(131072, 288, 96), (131072, 384, 96), (131072, 96, 384), (32768, 576, 192), (2048, 2304, 768), (2048, 3072, 768), (2048, 768, 3072)]:
    dy = bf(torch.randn(M, N, device="cuda")); x = bf(torch.randn(M, K, device="cuda"))
    want = dy.float().t() @ x.float()
    for impl, name in ((1, "mma"), (2, "tc05")):
        dW, db = ops.linear_wgrad(dy, x, impl=impl)
        rel = ((dW - want).norm() / want.norm()).item(); relb = ((db - dy.float().sum(0)).norm() / dy.float().sum(0).norm()).item()
        for _ in range(2): ops.linear_wgrad(dy, x, impl=impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib_ms = []
        torch.cuda.synchronize()
        import time
        from tulip_b200._lib import load_library, ptr, current_stream
        lib = load_library()
        dWb = torch.zeros(N, K, device="cuda"); dbb = torch.zeros(N, device="cuda")
        e0.record()
        for _ in range(20): lib.tulip_gemm_tn(ptr(dy), ptr(x), ptr(dWb), ptr(dbb), M, N, K, impl, current_stream())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"TN M={M} N={N} K={K} {name}: rel {rel:.2e} db {relb:.2e}  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s  {(M*K+M*N)*2/ms/1e6:.0f} GB/s", flush=True)
