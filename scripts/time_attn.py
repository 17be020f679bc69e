"""Stand-alone timing of the attention core per stage (CUDA events, L2 flushed between iterations)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
from tulip_b200._lib import load_library, ptr, current_stream
lib = load_library()
B = 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for s in range(4):
    H, W, Cc = 16 >> s, 256 >> s, 96 << s
    heads = Cc // 32; T = B * H * W
    qkv = torch.randn(T, 3 * Cc, device="cuda").bfloat16(); dout = torch.randn(T, Cc, device="cuda").bfloat16()
    table = torch.randn(45, heads, device="cuda"); dqkv = torch.empty_like(qkv); dtab = torch.zeros_like(table)
    for nodb in (0, 1):
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.tulip_window_attention_bwd(ptr(qkv), ptr(table), ptr(dout), ptr(dqkv), None if nodb else ptr(dtab), B, H, W, Cc, heads, 2, 8, 1, 4, 1, 2, 8, current_stream())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"stage {s} bwd nodb={nodb}: min {min(ts[1:]):.1f} us  {[round(t,1) for t in ts]}")
