"""Per-kernel parity: every hand-written CUDA kernel, called through the C ABI, against the fp32 oracle on the
same bf16-exact inputs.

Tolerances (SURVEY.md 8c -- the north star's 1e-3 is attainable per kernel, not end-to-end in bf16):
  * kernels whose only rounding is the final bf16 store: rel-L2 <= 1e-3 against the bf16-rounded fp32 oracle and
    >= 99 % of elements within 1 bf16 ulp;
  * the attention core keeps P and dS as two-term bf16 splits (hi + lo) for their MMAs: rel-L2 <= 1e-3 against the bf16-rounded
    fp32 oracle, like a single-rounding kernel; the qkv-GEMM -> attention -> proj-GEMM chain checked against the reference
    WindowAttention fixtures stores bf16 three times on inputs of std ~2: rel-L2 <= 8e-3 (measured 4.8-5.0e-3);
  * fp32 outputs (weight / bias / LayerNorm / table gradients, loss): rel-L2 <= 1e-3 (2e-3 where bf16 operands feed them).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tulip_oracle as O
from oracle.params import TULIP_BASE
from tests.util import bf16r, dec, load_modules, mod_params, rel_l2, ulp_frac

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from tulip_b200 import ops
    return ops


@pytest.fixture(scope="module")
def mods(golden_dir):
    return load_modules(golden_dir)


def rnd(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return bf16r(torch.randn(*shape, generator=gen) * scale)


IMPLS = [0]         # one GEMM implementation (tcgen05); the warp-MMA backend of round 1 is gone


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("M,N,K", [(256, 96, 96), (1000, 288, 96), (128, 384, 96), (64, 2304, 768), (4096, 96, 384), (333, 192, 1536)])
def test_gemm_nt_store_bias(ops, impl, M, N, K):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    want = F.linear(x, w, b)
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), impl=impl).float().cpu()
    assert rel_l2(got, bf16r(want)) <= 1e-3 and ulp_frac(got, bf16r(want)) >= 0.99


@pytest.mark.parametrize("impl", IMPLS)
def test_gemm_nt_gelu_resid_dgelu(ops, impl):
    M, N, K = 512, 384, 96
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    pre = F.linear(x, w, b)
    act, pre_got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_GELU, impl=impl)
    assert rel_l2(pre_got.float().cpu(), bf16r(pre)) <= 1e-3
    assert rel_l2(act.float().cpu(), bf16r(F.gelu(pre))) <= 1e-3               # exact-erf GELU (tulip.py:183)
    res = rnd(M, N, seed=4)
    scale = torch.tensor([0.0, 1.0 / 0.9, 1.0, 1.0 / 0.95])                     # DropPath scales, 128 rows per sample
    want = res + scale.repeat_interleave(128)[:, None] * pre
    got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_RESID, aux=res.cuda(), row_scale=scale.cuda(),
                     rows_per_sample=128, impl=impl).float().cpu()
    assert rel_l2(got, bf16r(want)) <= 1e-3
    # backward through GELU: out = (x w^T) * gelu'(aux)
    aux = rnd(M, N, seed=5)
    a = aux.clone().requires_grad_(True)
    F.gelu(a).sum().backward()
    want = F.linear(x, w) * a.grad
    got = ops.linear(x.cuda(), w.cuda(), None, epilogue=ops.EPI_DGELU, aux=aux.cuda(), impl=impl).float().cpu()
    assert rel_l2(got, bf16r(want)) <= 1e-3


@pytest.mark.parametrize("M,N,K,epi", [(76001, 288, 96, "store"), (76001, 96, 384, "resid"), (76001, 384, 96, "gelu"),
                                         (38000, 576, 192, "store"), (76001, 96, 288, "dgelu")])
def test_gemm_nt_resident_weight_schedule(ops, M, N, K, epi):
    """Shapes of the wide stages at full batch: enough 128-row panels per CTA for the B-stationary schedule (weights
    resident in shared memory, A blocks shared by the tiles of a panel), ragged M, every TMA-store epilogue."""
    x, w, b = rnd(M, K, seed=11), rnd(N, K, seed=12, scale=K ** -0.5), rnd(N, seed=13)
    pre = F.linear(x, w, b)
    if epi == "store":
        got = ops.linear(x.cuda(), w.cuda(), b.cuda()).float().cpu()
        want = pre
    elif epi == "gelu":
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_GELU, save_pre=False)[0].float().cpu()   # wide tile
        got2 = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_GELU)                                # narrow, two outputs
        want = F.gelu(pre)
        assert torch.equal(got2[0].float().cpu(), got) and rel_l2(got2[1].float().cpu(), bf16r(pre)) <= 1e-3
    elif epi == "resid":
        res = rnd(M, N, seed=14)
        rows_per = 1000
        scale = (torch.arange(-(-M // rows_per)) % 3).float() * 0.55
        want = res + scale.repeat_interleave(rows_per)[:M, None] * pre
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_RESID, aux=res.cuda(), row_scale=scale.cuda(),
                         rows_per_sample=rows_per).float().cpu()
    else:
        aux = rnd(M, N, seed=15)
        a = aux.clone().requires_grad_(True)
        F.gelu(a).sum().backward()
        want = F.linear(x, w) * a.grad
        got = ops.linear(x.cuda(), w.cuda(), None, epilogue=ops.EPI_DGELU, aux=aux.cuda()).float().cpu()
    assert rel_l2(got, bf16r(want)) <= 1e-3 and ulp_frac(got, bf16r(want)) >= 0.99


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("M,N,K", [(4096, 288, 96), (1000, 96, 384), (64, 768, 3072), (8192, 1536, 96)])
def test_gemm_tn_wgrad(ops, impl, M, N, K):
    dy, x = rnd(M, N, seed=1), rnd(M, K, seed=2)
    dW, db = ops.linear_wgrad(dy.cuda(), x.cuda(), impl=impl)
    assert rel_l2(dW.cpu(), dy.double().T @ x.double()) <= 1e-4
    assert rel_l2(db.cpu(), dy.double().sum(0)) <= 1e-4


@pytest.mark.parametrize("rows,C", [(1000, 96), (512, 192), (300, 384), (130, 768), (40, 1536), (24, 3072)])
def test_layernorm_fwd_bwd(ops, rows, C):
    x, w, b = rnd(rows, C, seed=1), bf16r(1 + 0.1 * torch.randn(C)), bf16r(0.1 * torch.randn(C))
    xr = x.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.layer_norm(xr, (C,), wr, br, 1e-6)
    dy, dres = rnd(rows, C, seed=2), rnd(rows, C, seed=3)
    y.backward(dy)
    got, stats = ops.layernorm(x.cuda(), w.cuda(), b.cuda())
    assert rel_l2(got.float().cpu(), bf16r(y)) <= 1e-3 and ulp_frac(got.float().cpu(), bf16r(y)) >= 0.99
    assert rel_l2(stats[:, 0].cpu(), x.mean(1)) <= 1e-4
    dx, dw, db = ops.layernorm_bwd(x.cuda(), w.cuda(), stats, dy.cuda(), dres.cuda())
    assert rel_l2(dx.float().cpu(), bf16r(xr.grad + dres)) <= 1e-3
    assert rel_l2(dw.cpu(), wr.grad) <= 1e-3 and rel_l2(db.cpu(), br.grad) <= 1e-3


def test_layernorm_merge_gather_fwd_bwd(ops, mods):
    """PatchMerging = gather + LN + reduction (tulip.py:101-106) against the reference-generated fixture."""
    p = {k: v for k, v in mod_params(mods, "merge").items()}
    x = dec(mods, "merge.x")                                                     # (2,4,8,96)
    B, H, W, C = x.shape
    xn, stats = ops.layernorm(x.cuda(), p["norm.weight"].cuda(), p["norm.bias"].cuda(), merge=(B, H, W))
    y = ops.linear(xn, p["reduction.weight"].cuda(), None).float().cpu().view(B, H // 2, W // 2, 2 * C)
    assert rel_l2(y, mods["merge.y"]) <= 3e-3
    # backward of the gather+LN against autograd on the oracle
    xr = x.clone().requires_grad_(True)
    xm = torch.cat([xr[:, 0::2, 0::2], xr[:, 1::2, 0::2], xr[:, 0::2, 1::2], xr[:, 1::2, 1::2]], -1)
    yn = F.layer_norm(xm, (4 * C,), p["norm.weight"], p["norm.bias"], 1e-6)
    dy = rnd(*yn.shape, seed=5)
    yn.backward(dy)
    dx, dw, db = ops.layernorm_bwd(x.cuda(), p["norm.weight"].cuda(), stats, dy.view(-1, 4 * C).cuda(), None, merge=(B, H, W))
    assert rel_l2(dx.float().cpu().view_as(x), bf16r(xr.grad)) <= 1e-3


@pytest.mark.parametrize("shift", [0, 1])
def test_window_attention_fwd_bwd_fixture(ops, mods, shift):
    """Attention core + qkv/proj GEMMs against the reference WindowAttention fixture (fwd and bwd)."""
    tag = f"attn_shift{shift}"
    p = mod_params(mods, tag)
    x, gy = dec(mods, f"{tag}.x"), dec(mods, f"{tag}.gy")
    B, H, W, C = x.shape
    sh = (1, 4) if shift else (0, 0)
    qkv = ops.linear(x.view(-1, C).cuda(), p["qkv.weight"].cuda(), p["qkv.bias"].cuda())
    o = ops.window_attention(qkv, p["relative_position_bias_table"].cuda(), B, H, W, 3, (2, 8), sh, bool(shift))
    y = ops.linear(o, p["proj.weight"].cuda(), p["proj.bias"].cuda()).float().cpu().view(B, H, W, C)
    assert rel_l2(y, mods[f"{tag}.y"]) <= 8e-3
    do = ops.linear(gy.view(-1, C).cuda(), p["proj.weight"].t().contiguous().cuda(), None)
    dqkv, dtab = ops.window_attention_bwd(qkv, p["relative_position_bias_table"].cuda(), do, B, H, W, 3, (2, 8), sh, bool(shift))
    dx = ops.linear(dqkv, p["qkv.weight"].t().contiguous().cuda(), None).float().cpu().view(B, H, W, C)
    assert rel_l2(dx, mods[f"{tag}.gx"]) <= 1e-2
    assert rel_l2(dtab.cpu(), mods[f"{tag}.g_table"]) <= 1e-2
    dWqkv, _ = ops.linear_wgrad(dqkv, x.view(-1, C).cuda())
    assert rel_l2(dWqkv.cpu(), mods[f"{tag}.g_qkv_w"]) <= 1e-2


def test_window_attention_backup_window(ops, mods):
    p = mod_params(mods, "attn_backup")
    x = dec(mods, "attn_backup.x")                                               # (2,1,32,96): H < win_h
    B, H, W, C = x.shape
    qkv = ops.linear(x.view(-1, C).cuda(), p["qkv.weight"].cuda(), p["qkv.bias"].cuda())
    o = ops.window_attention(qkv, p["relative_position_bias_table"].cuda(), B, H, W, 3, (1, 16), (0, 8), True, (2, 8))
    y = ops.linear(o, p["proj.weight"].cuda(), p["proj.bias"].cuda()).float().cpu().view(B, H, W, C)
    assert rel_l2(y, mods["attn_backup.y"]) <= 8e-3


@pytest.mark.parametrize("B,H,W,C,heads", [(2, 4, 64, 384, 12), (3, 2, 32, 768, 24), (1, 16, 256, 96, 3)])
def test_window_attention_core_vs_oracle(ops, B, H, W, C, heads):
    """The attention core alone (given qkv) for the wider stages, shifted, vs the oracle math."""
    qkv = rnd(B * H * W, 3 * C, seed=1)
    table = bf16r(0.2 * torch.randn(45, heads))
    q = qkv.view(B, H, W, 3 * C)
    q = torch.roll(q, (-1, -4), (1, 2))
    qw = O.window_partition(q, (2, 8))
    Bn = qw.shape[0]
    t = qw.view(Bn, 16, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = (t[0] * 32 ** -0.5) @ t[1].transpose(-2, -1)
    from oracle.index_ops import relative_position_index
    attn = attn + O.relative_position_bias(table, torch.from_numpy(relative_position_index((2, 8)))).unsqueeze(0)
    mask = O.shift_mask(H, W, (2, 8), (1, 4))
    attn = (attn.view(B, -1, heads, 16, 16) + mask[None, :, None]).view(Bn, heads, 16, 16).softmax(-1)
    o = (attn @ t[2]).permute(0, 2, 1, 3).reshape(Bn, 16, C)
    want = torch.roll(O.window_reverse(o, (2, 8), B, H, W), (1, 4), (1, 2)).reshape(-1, C)
    got = ops.window_attention(qkv.cuda(), table.cuda(), B, H, W, heads, (2, 8), (1, 4), True).float().cpu()
    # P enters its MMA as a two-term bf16 split, so the only rounding left is the bf16 store of the output
    assert rel_l2(got, bf16r(want)) <= 1e-3
    assert rel_l2(got, want) <= 2.5e-3


def test_patch_embed_fwd_bwd(ops, mods):
    p = mod_params(mods, "embed")
    x = dec(mods, "embed.x")                                                     # (2,1,4,64)
    y = ops.patch_embed(x.cuda(), p["proj.weight"].cuda(), p["proj.bias"].cuda(), p["norm.weight"].cuda(), p["norm.bias"].cuda())
    assert rel_l2(y.float().cpu(), bf16r(torch.from_numpy(mods["embed.y"]))) <= 1e-3
    pp = {f"patch_embed.{k}": v.clone().requires_grad_(True) for k, v in p.items()}
    yo = O.patch_embed(x, pp, TULIP_BASE)
    dy = rnd(*yo.shape, seed=3)
    yo.backward(dy)
    dw, db, dlw, dlb = ops.patch_embed_bwd(x.cuda(), p["proj.weight"].cuda(), p["proj.bias"].cuda(), p["norm.weight"].cuda(), dy.cuda())
    assert rel_l2(dw.cpu(), pp["patch_embed.proj.weight"].grad) <= 1e-4
    assert rel_l2(db.cpu(), pp["patch_embed.proj.bias"].grad) <= 1e-4
    assert rel_l2(dlw.cpu(), pp["patch_embed.norm.weight"].grad) <= 1e-4
    assert rel_l2(dlb.cpu(), pp["patch_embed.norm.bias"].grad) <= 1e-4


def test_l1_loss(ops):
    pred, y = torch.rand(2, 1, 64, 1024), torch.rand(2, 1, 64, 1024)
    y[0, 0, :4] = pred[0, 0, :4]                                                # exact zeros: sign(0) = 0 region
    loss, pixel = ops.l1_loss(pred.cuda(), y.cuda(), True)
    assert abs(loss.item() - (pred - y).abs().mean().item()) <= 1e-6
    assert abs(pixel.item() - (torch.expm1(pred) - torch.expm1(y)).abs().mean().item()) <= 2e-6


@pytest.mark.parametrize("M,N,K,epi", [
    (8192, 384, 1536, "resid"),      # fc2, stage 2: BN = 96 pairs
    (2048, 3072, 768, "gelu"),       # fc1, stage 3: BN = 192 pairs
    (2048, 2304, 768, "store"),      # qkv, stage 3
    (2048, 768, 3072, "resid"),      # fc2, stage 3: 48 K blocks per tile
    (1152, 768, 768, "store"),       # 9 row tiles: the last pair has a phantom tile
    (300, 192, 1536, "resid"),       # ragged rows inside the last real tile
    (32768, 192, 768, "store"),      # dX of fc1, stage 1
])
def test_gemm_nt_cta_pairs(ops, M, N, K, epi):
    """Deep-K launches run as CTA pairs (cta_group::2, schedule 2 of tulip_gemm_nt_plan): same tolerances as every
    single-rounding GEMM epilogue, and the plan says pairs were used."""
    import ctypes as C
    code = {"store": ops.EPI_STORE, "gelu": ops.EPI_GELU, "resid": ops.EPI_RESID}[epi]
    out = (C.c_int * 10)()
    lib = ops.load_library()
    prev = lib.tulip_gemm_nt_pairs_mode(1)                     # opt-in schedule (tulip_b200.h): switched on for this test only
    try:
        assert lib.tulip_gemm_nt_plan(M, N, K, code, 0, out) == 0 and out[1] == 2
        _run_pairs_case(ops, M, N, K, epi)
    finally:
        lib.tulip_gemm_nt_pairs_mode(max(prev, 0))


def _run_pairs_case(ops, M, N, K, epi):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    pre = F.linear(x, w, b)
    if epi == "store":
        want = pre
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), impl=2)
    elif epi == "gelu":
        want = F.gelu(pre)
        got, _ = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_GELU, impl=2, save_pre=False)
    else:
        res, rows_per = rnd(M, N, seed=4), 4
        scale = torch.rand(-(-M // rows_per), generator=torch.Generator().manual_seed(5)) * 2
        want = res + scale.repeat_interleave(rows_per)[:M, None] * pre
        got = ops.linear(x.cuda(), w.cuda(), b.cuda(), epilogue=ops.EPI_RESID, aux=res.cuda(), row_scale=scale.cuda(),
                         rows_per_sample=rows_per, impl=2)
    got = got.float().cpu()
    assert rel_l2(got, bf16r(want)) <= 1e-3 and ulp_frac(got, bf16r(want)) >= 0.99
