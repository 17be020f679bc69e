"""A few launches of the fused W-MSA half-block at BASELINE cfg2 stage-0 size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
from tests.test_gpu_wmsa import make_block_params, rnd
B, H, W, C, heads = 32, 16, 256, 96, 3
p = {k: v.cuda() for k, v in make_block_params(C, heads, 1).items()}
x = rnd(B * H * W, C, seed=2, scale=1.5).cuda().to(torch.bfloat16)
wq, wp = p["attn.qkv.weight"].to(torch.bfloat16), p["attn.proj.weight"].to(torch.bfloat16)
for _ in range(4):
    ops.wmsa_block(x, p["norm1.weight"], p["norm1.bias"], wq, p["attn.qkv.bias"], wp, p["attn.proj.bias"],
                   p["attn.relative_position_bias_table"], B, H, W, heads, (2, 8), (1, 4), True)
torch.cuda.synchronize()
