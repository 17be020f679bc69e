"""Weight-gradient GEMMs of a half-block, stand-alone kernel (one launch per problem) against the grouped persistent launch:
back-to-back launches timed with one event pair (no per-launch event overhead).  usage: time_tn.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REP = 40


def timed(fn):
    # the launches are replayed from a CUDA graph, as in the training step: host-side launch cost (tensor-map encoding, the
    # work-item plan) is not part of the reading
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


print(f"| stage | group | GFLOP | MB | singles us | grouped us | singles TF/s | grouped TF/s | grouped GB/s |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
for s, (T, C) in enumerate([(B * 4096, 96), (B * 1024, 192), (B * 256, 384), (B * 64, 768)]):
    for name, shapes in (("fc2+fc1", [(T, C, 4 * C), (T, 4 * C, C)]), ("proj+qkv", [(T, C, C), (T, 3 * C, C)])):
        probs = []
        for (M, N, K) in shapes:
            dy = torch.randn(M, N, device="cuda").bfloat16()
            x = torch.randn(M, K, device="cuda").bfloat16()
            dW = torch.zeros(N, K, device="cuda")
            db = torch.zeros(N, device="cuda")
            probs.append(dict(dY=dy, ldy=N, X=x, ldx=K, K1=K, M=M, N=N, K=K, dW=dW, lddw=K, db=db))
        lib = ops.load_library()

        def singles():
            for p in probs:
                ops.check(lib.tulip_gemm_tn(ops.ptr(p["dY"]), ops.ptr(p["X"]), ops.ptr(p["dW"]), ops.ptr(p["db"]), p["M"], p["N"], p["K"], 2,
                                            ops.current_stream()), "tulip_gemm_tn")

        def grouped():
            ops.gemm_tn_group(probs)

        fl = sum(2.0 * M * N * K for M, N, K in shapes)
        by = sum(2.0 * (M * N + M * K) + 4.0 * N * K for M, N, K in shapes)
        t1, t2 = timed(singles), timed(grouped)
        print(f"| {s} | {name} | {fl / 1e9:.2f} | {by / 1e6:.1f} | {t1:.1f} | {t2:.1f} | {fl / t1 / 1e6:.0f} | {fl / t2 / 1e6:.0f} | {by / t2 / 1e3:.0f} |")
