"""Fused MLP half-block kernel (tulip_mlp_block_fwd, csrc/mlp.cu) against
  * the oracle's MLP half (oracle/tulip_oracle.py: layer_norm + mlp, restating tulip.py:347-352, 194-200) on seeded shapes,
    with DropPath scales and ragged token counts (partial last tile);
  * the unfused kernel chain (LayerNorm -> fc1 + GELU GEMM -> fc2 GEMM + residual), which rounds at the same points;
  * its training by-products (LayerNorm output, statistics, activated hidden tensor) against the kernels they replace;
  * the reference SwinTransformerBlock fixture (tests/golden/modules.npz `block.*`) together with the fused attention half.

Tolerance: y = x + branch is stored once as bf16; the branch is computed from bf16 operands rounded at the LayerNorm output and at
the activated hidden tensor (fp32 accumulation, fp32 GELU / LayerNorm statistics).  Asserted: rel-L2 <= 2e-3 against the
bf16-rounded fp32 oracle and <= 5e-4 against the unfused chain."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tulip_oracle as O
from tests.util import bf16r, dec, load_modules, mod_params, rel_l2, ulp_frac

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return bf16r(torch.randn(*shape, generator=gen) * scale)


def make_params(C, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: bf16r(torch.randn(*s, generator=g) * sc)
    return {"norm2.weight": bf16r(1.0 + 0.1 * torch.randn(C, generator=g)), "norm2.bias": r(C, sc=0.05),
            "mlp.fc1.weight": r(4 * C, C, sc=0.7 * C ** -0.5), "mlp.fc1.bias": r(4 * C, sc=0.05),
            "mlp.fc2.weight": r(C, 4 * C, sc=0.7 * (4 * C) ** -0.5), "mlp.fc2.bias": r(C, sc=0.05)}


def fused(ops, x, p, scales=None, rps=1, save=False):
    out = ops.mlp_block(x.cuda(), p["norm2.weight"].cuda(), p["norm2.bias"].cuda(), p["mlp.fc1.weight"].cuda(), p["mlp.fc1.bias"].cuda(),
                        p["mlp.fc2.weight"].cuda(), p["mlp.fc2.bias"].cuda(), None if scales is None else scales.cuda(), rps, save)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("T,rps,drop", [(128, 128, False), (200, 50, True), (1000, 250, False), (8192, 4096, True), (131072 + 40, 1, False)])
def test_mlp_block_vs_oracle(T, rps, drop):
    from tulip_b200 import ops
    C = 96
    p = make_params(C, seed=T % 97)
    x = rnd(T, C, seed=3, scale=1.5)
    scales = None
    if drop:
        nb = (T + rps - 1) // rps
        scales = torch.tensor([0.0 if b % 3 == 1 else 1.0 / 0.9 for b in range(nb)], dtype=torch.float32)
    xn = O.layer_norm(x, p["norm2.weight"], p["norm2.bias"], 1e-6)
    br = O.mlp(xn, {f"b.{k}": v for k, v in p.items()}, "b.mlp")
    sc = torch.ones(T) if scales is None else scales[torch.arange(T) // rps]
    want = x + br * sc[:, None]
    got = fused(ops, x, p, scales, rps).float().cpu()
    e = rel_l2(got, bf16r(want))
    print(f"\n[mlp T={T}] y rel-L2 {e:.3e}")
    assert e <= 2e-3
    if scales is not None:
        assert torch.equal(got[sc == 0], x[sc == 0])


def test_mlp_block_matches_unfused_chain_and_byproducts():
    from tulip_b200 import ops
    T, C = 4096 + 77, 96
    p = make_params(C, seed=11)
    x = rnd(T, C, seed=5, scale=1.5)
    y, xn, stats, hact = fused(ops, x, p, save=True)
    xc = x.cuda()
    xn_u, st_u = ops.layernorm(xc, p["norm2.weight"].cuda(), p["norm2.bias"].cuda())
    h_u, _ = ops.linear(xn_u, p["mlp.fc1.weight"].cuda(), p["mlp.fc1.bias"].cuda(), epilogue=ops.EPI_GELU, save_pre=False)
    y_u = ops.linear(h_u, p["mlp.fc2.weight"].cuda(), p["mlp.fc2.bias"].cuda(), epilogue=ops.EPI_RESID, aux=xc)
    assert rel_l2(xn.float(), xn_u.float()) <= 2e-4 and ulp_frac(xn.float(), xn_u.float()) >= 0.999
    assert rel_l2(stats, st_u) <= 1e-5
    assert rel_l2(hact.float(), h_u.float()) <= 5e-4
    e = rel_l2(y.float(), y_u.float())
    print(f"\n[mlp vs unfused chain] y rel-L2 {e:.3e}")
    assert e <= 5e-4
    y2 = fused(ops, x, p, save=False)                        # inference mode: same y, nothing else written
    assert torch.equal(y2, y)


def test_fused_block_reference_fixture(golden_dir):
    """Reference SwinTransformerBlock(96, 3, window (2,8), shift=True): fused attention half then fused MLP half."""
    from tulip_b200 import ops
    mods = load_modules(golden_dir)
    p = mod_params(mods, "block")
    x = dec(mods, "block.x")
    B, H, W, C = x.shape
    xm = ops.wmsa_block(x.reshape(-1, C).cuda(), p["norm1.weight"].cuda(), p["norm1.bias"].cuda(), p["attn.qkv.weight"].cuda(),
                        p["attn.qkv.bias"].cuda(), p["attn.proj.weight"].cuda(), p["attn.proj.bias"].cuda(),
                        p["attn.relative_position_bias_table"].cuda(), B, H, W, 3, (2, 8), (1, 4), True)
    y = fused(ops, xm.float().cpu(), p)
    e = rel_l2(y.float().cpu().view(B, H, W, C), bf16r(torch.from_numpy(mods["block.y"])))
    print(f"\n[fused wmsa + fused mlp vs reference block fixture] rel-L2 {e:.3e}")
    assert e <= 8e-3


def test_mlp_block_unsupported_shape_raises():
    from tulip_b200 import ops
    from tulip_b200._lib import load_library
    lib = load_library()
    assert lib.tulip_mlp_block_supported(131072, 96) == 1 and lib.tulip_mlp_block_supported(32768, 192) == 0
    with pytest.raises(RuntimeError):
        fused(ops, rnd(128, 192), make_params(192, 1))
