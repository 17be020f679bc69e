"""Runs one NT GEMM shape a few times (target for ncu --set full captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tulip_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
epi = int(sys.argv[4]) if len(sys.argv) > 4 else 0
x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16(); b = torch.randn(N, device="cuda")
aux = torch.randn(M, N, device="cuda").bfloat16()
for _ in range(4):
    if epi == 0: ops.linear(x, w, b, impl=2)
    elif epi == 1: ops.linear(x, w, b, epilogue=1, impl=2, save_pre=False)   # the executor's configuration (pre-activation recomputed)
    elif epi == 2: ops.linear(x, w, b, epilogue=2, aux=aux, impl=2)
    elif epi == 5: ops.linear(x, w, None, epilogue=5, aux=aux, impl=2)
    elif epi == 9:
        ops.linear_wgrad(aux, x, impl=2)
torch.cuda.synchronize()
