"""Runs the window-attention core (fwd and bwd) of one stage a few times (target for ncu captures).
usage: one_attn.py stage [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
s = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H, W, C = 16 >> s, 256 >> s, 96 << s
heads = C // 32
T = B * H * W
qkv = torch.randn(T, 3 * C, device="cuda").bfloat16()
dout = torch.randn(T, C, device="cuda").bfloat16()
table = torch.randn(45, heads, device="cuda")
for _ in range(4):
    ops.window_attention(qkv, table, B, H, W, heads, shift=(1, 4), masked=True)
    ops.window_attention_bwd(qkv, table, dout, B, H, W, heads, shift=(1, 4), masked=True)
torch.cuda.synchronize()
