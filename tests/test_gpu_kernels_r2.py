"""Per-kernel GPU parity of the scatter / gather / split / head epilogues against fixtures computed by the reference's own
modules (tests/golden/modules_r2.npz, oracle/make_golden_r2.py): PatchUnmerging (tulip.py:109-123), the skip Linear on
cat([x, skip]) (:715-716) and norm_up + PixelShuffleHead + decoder_pred + L1 (:720-731, 690-693), forward and backward.

Tolerances.  A kernel that rounds once (fp32 accumulate -> one bf16 store) is held to rel-L2 <= 1e-3 against the bf16-rounded
reference output.  fp32 outputs (pred, weight gradients accumulated in fp32 from exact bf16 operands) to 1e-3 against the
reference directly.  Chains of two rounded stages (dh is stored as bf16 before it feeds the next GEMM; LayerNorm output is stored
as bf16 before the head GEMM) are held to 4e-3: two independent bf16 roundings of 2^-9 relative each, measured ~2e-3."""
import os

import numpy as np
import pytest
import torch

from tests.util import bf16r, dec, rel_l2
from oracle.params import bf16_bits_to_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(golden_dir):
    return np.load(os.path.join(golden_dir, "modules_r2.npz"))


def P(mods, key):
    v = mods[key]
    return torch.from_numpy(bf16_bits_to_f32(v) if v.dtype == np.uint16 else v).cuda()


def bf(t):
    return t.to(torch.bfloat16).contiguous()


def test_patch_unmerging_fwd_bwd(mods):
    from tulip_b200 import ops
    x = P(mods, "unmerge.x")                                   # [B, H, W, C]
    B, H, W, C = x.shape
    Cc, T = C // 2, B * H * W
    w = P(mods, "unmerge.p.expand.weight").reshape(2 * C, C)
    b = P(mods, "unmerge.p.expand.bias")
    wp, bp = ops.permute_rows_for_shuffle(w, 4, Cc), ops.permute_rows_for_shuffle(b, 4, Cc)
    xb, wb = bf(x.reshape(T, C)), bf(wp)
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_PIXSHUF, A=xb, lda=C, K1=C, B=wb, ldb=C, M=T, N=2 * C, K=C, bias=bp.contiguous(), out=out, ldo=Cc,
                   g_H=H, g_W=W, g_Cc=Cc)
    y = P(mods, "unmerge.y")
    assert rel_l2(out.float(), bf16r(y)) <= 1e-3
    # backward: dX = gather(gy) . W' (A_UNSHUFFLE), dW = gather(gy)^T . x written back in the reference's row order
    gy = bf(P(mods, "unmerge.gy"))
    wtb = bf(wp.t())                                           # [C, 2C]
    gx = torch.empty((T, C), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_STORE, A=gy, lda=Cc, K1=2 * C, B=wtb, ldb=2 * C, M=T, N=C, K=2 * C, out=gx, ldo=C, a_mode=ops.A_UNSHUFFLE,
                   g_H=H, g_W=W, g_Cc=Cc)
    assert rel_l2(gx.float().reshape(B, H, W, C), bf16r(P(mods, "unmerge.gx"))) <= 1e-3
    dW = torch.zeros((2 * C, C), dtype=torch.float32, device="cuda")
    db = torch.zeros(2 * C, dtype=torch.float32, device="cuda")
    ops.gemm_tn_ex(dY=gy, ldy=Cc, X=xb, ldx=C, K1=C, M=T, N=2 * C, K=C, y_mode=ops.A_UNSHUFFLE, g_H=H, g_W=W, g_Cc=Cc, dW=dW, lddw=C,
                   db=db, perm_R2=4, perm_Cc=Cc)
    assert rel_l2(dW, P(mods, "unmerge.g_w").reshape(2 * C, C)) <= 1e-3
    assert rel_l2(db, P(mods, "unmerge.g_b")) <= 1e-3


def test_skip_linear_fwd_bwd(mods):
    from tulip_b200 import ops
    x, skip = P(mods, "skip.x"), P(mods, "skip.skip")
    B, H, W, C = x.shape
    T = B * H * W
    w, b = P(mods, "skip.p.weight"), P(mods, "skip.p.bias")    # [C, 2C]
    xb, sb, wb = bf(x.reshape(T, C)), bf(skip.reshape(T, C)), bf(w)
    out = torch.empty((T, C), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_STORE, A=xb, lda=C, A2=sb, lda2=C, K1=C, B=wb, ldb=2 * C, M=T, N=C, K=2 * C, bias=b.contiguous(), out=out, ldo=C)
    assert rel_l2(out.float().reshape(B, H, W, C), bf16r(P(mods, "skip.y"))) <= 1e-3
    gy = bf(P(mods, "skip.gy").reshape(T, C))
    wtb = bf(w.t())                                            # [2C, C]
    gx = torch.empty((T, C), dtype=torch.bfloat16, device="cuda")
    gs = torch.empty((T, C), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_SPLIT2, A=gy, lda=C, K1=C, B=wtb, ldb=C, M=T, N=2 * C, K=C, out=gx, ldo=C, out2=gs, ldo2=C, split_col=C)
    assert rel_l2(gx.float().reshape(B, H, W, C), bf16r(P(mods, "skip.gx"))) <= 1e-3
    assert rel_l2(gs.float().reshape(B, H, W, C), bf16r(P(mods, "skip.gskip"))) <= 1e-3
    dW = torch.zeros((C, 2 * C), dtype=torch.float32, device="cuda")
    db = torch.zeros(C, dtype=torch.float32, device="cuda")
    ops.gemm_tn_ex(dY=gy, ldy=C, X=xb, ldx=C, X2=sb, ldx2=C, K1=C, M=T, N=C, K=2 * C, dW=dW, lddw=2 * C, db=db)
    assert rel_l2(dW, P(mods, "skip.g_w")) <= 1e-3
    assert rel_l2(db, P(mods, "skip.g_b")) <= 1e-3


@pytest.mark.parametrize("E", [96, 192])
def test_head_fwd_bwd(mods, E):
    """norm_up -> conv_expand -> LeakyReLU -> PixelShuffle(4) -> decoder_pred -> L1, and its backward."""
    from tulip_b200 import ops
    tag, r = f"head{E}", 4
    x = P(mods, f"{tag}.x")                                    # [B, H, W, E] NHWC tokens
    B, H, W, _ = x.shape
    T, N = B * H * W, E * r * r
    nw, nb = P(mods, f"{tag}.norm.weight"), P(mods, f"{tag}.norm.bias")
    we = P(mods, f"{tag}.ps.conv_expand.0.weight").reshape(N, E)
    be = P(mods, f"{tag}.ps.conv_expand.0.bias")
    wd = P(mods, f"{tag}.dec.weight").reshape(E).contiguous()
    target = P(mods, f"{tag}.target").contiguous()
    wep, bep = ops.permute_rows_for_shuffle(we, r * r, E), ops.permute_rows_for_shuffle(be, r * r, E).contiguous()
    xb = bf(x.reshape(T, E))
    xn, stats = ops.layernorm(xb, nw, nb)
    pred = torch.zeros((B, 1, H * r, W * r), dtype=torch.float32, device="cuda")
    wb = bf(wep)
    ops.gemm_nt_ex(ops.EPI_HEAD, A=xn, lda=E, K1=E, B=wb, ldb=E, M=T, N=N, K=E, bias=bep, wd=wd, pred=pred, hd_H=H, hd_W=W, hd_r=r, hd_E=E)
    want = P(mods, f"{tag}.pred")
    assert rel_l2(pred, want) <= 4e-3                          # LayerNorm output is a bf16 tensor: two rounded stages
    loss, _ = ops.l1_loss(pred, target, log_transform=False)
    assert abs(loss.item() - float(mods[f"{tag}.loss"])) <= 2e-3 * float(mods[f"{tag}.loss"])
    # backward.  dpred = sign(pred - target) / numel is formed inside the kernel from pred / target.
    dh = torch.empty((T, N), dtype=torch.bfloat16, device="cuda")
    dwd = torch.zeros(E, dtype=torch.float32, device="cuda")
    gscale = torch.ones(1, dtype=torch.float32, device="cuda")
    ops.gemm_nt_ex(ops.EPI_HEAD_BWD, A=xn, lda=E, K1=E, B=wb, ldb=E, M=T, N=N, K=E, bias=bep, wd=wd, pred=pred, target=target,
                   gscale=gscale, dwd=dwd, out=dh, ldo=N, hd_H=H, hd_W=W, hd_r=r, hd_E=E)
    dWe = torch.zeros((N, E), dtype=torch.float32, device="cuda")
    dbe = torch.zeros(N, dtype=torch.float32, device="cuda")
    ops.gemm_tn_ex(dY=dh, ldy=N, X=xn, ldx=E, K1=E, M=T, N=N, K=E, dW=dWe, lddw=E, db=dbe, perm_R2=r * r, perm_Cc=E)
    dxn = torch.empty((T, E), dtype=torch.bfloat16, device="cuda")
    wtb = bf(wep.t())
    ops.gemm_nt_ex(ops.EPI_STORE, A=dh, lda=N, K1=N, B=wtb, ldb=N, M=T, N=E, K=N, out=dxn, ldo=E)
    if E == 96:
        # the executor's path at embed_dim 96: ONE launch (fused head backward) for dh, dxn and dwd -- same operands, same
        # roundings as the two GEMMs above (dh bf16, fp32 accumulation): everything within fp32 re-association
        lib = ops.load_library()
        assert lib.tulip_head_bwd_fused_supported(E, r) == 1
        dh2, dxn2 = torch.empty_like(dh), torch.empty_like(dxn)
        dwd2 = torch.zeros_like(dwd)
        ops.check(lib.tulip_head_bwd_fused(ops.ptr(xn), ops.ptr(dxn2), ops.ptr(wb), ops.ptr(wtb), ops.ptr(bep), ops.ptr(wd), ops.ptr(pred),
                                           ops.ptr(target), ops.ptr(gscale), ops.ptr(dh2), ops.ptr(dwd2), T, E, H, W, r,
                                           ops.current_stream()), "tulip_head_bwd_fused")
        assert rel_l2(dh2.float(), dh.float()) <= 1e-4          # (0.01 dp) wd against (dp wd) 0.01: one fp32 ulp before the bf16 store
        assert rel_l2(dwd2, dwd) <= 1e-5
        assert rel_l2(dxn2.float(), dxn.float()) <= 2e-3 and (dxn2.float() - dxn.float()).abs().max() <= 2 ** -7 * dxn.float().abs().max()
    gx, gnw, gnb = ops.layernorm_bwd(xb, nw, stats, dxn)
    # (1) kernel arithmetic: the same math in fp32 torch FROM THE SAME bf16 LayerNorm output and the kernel's own pred
    #     (LeakyReLU's slope jumps 100x at 0 and sign(pred - target) jumps at 0: both discontinuities are then evaluated on
    #     identical inputs, so the comparison sees rounding only)
    xn_f = xn.float().requires_grad_(True)
    we_f, be_f, wd_f = we.clone().requires_grad_(True), be.clone().requires_grad_(True), wd.clone().requires_grad_(True)
    pre = torch.nn.functional.linear(xn_f, we_f, be_f)                                          # [T, E r^2], reference row order c*16 + ij
    hsh = torch.nn.functional.pixel_shuffle(torch.nn.functional.leaky_relu(pre, 0.01).view(B, H, W, N).permute(0, 3, 1, 2), r)
    pred_t = (hsh * wd_f.view(1, E, 1, 1)).sum(1, keepdim=True)
    dpred = torch.sign(pred - target) / pred.numel()
    pred_t.backward(dpred)
    assert rel_l2(pred, pred_t) <= 1e-5
    assert rel_l2(dwd, wd_f.grad) <= 2e-3
    assert rel_l2(dWe, we_f.grad) <= 3e-3 and rel_l2(dbe, be_f.grad) <= 3e-3                   # dh is stored as bf16 first
    assert rel_l2(dxn.float(), xn_f.grad) <= 3e-3
    # (2) the reference fixture (fp32 LayerNorm output feeding the discontinuities): only T = 256 tokens are summed and ~0.2 % of
    #     the pre-activations change slope under the bf16 rounding of the LayerNorm output, which is a 5-10 % effect here
    g_we = P(mods, f"{tag}.g_we").reshape(-1, E)
    stride = N // g_we.shape[0]
    assert rel_l2(dwd, P(mods, f"{tag}.g_wd").reshape(E)) <= 2e-2
    assert rel_l2(dWe[::stride], g_we) <= 0.15 and rel_l2(dbe, P(mods, f"{tag}.g_be")) <= 0.15
    assert rel_l2(gx.float().reshape(B, H, W, E), P(mods, f"{tag}.gx")) <= 0.15
    assert rel_l2(gnw, P(mods, f"{tag}.g_norm_w")) <= 0.15 and rel_l2(gnb, P(mods, f"{tag}.g_norm_b")) <= 0.15


@pytest.mark.parametrize("M,C,K,res,scaled", [
    (1000, 96, 288, True, False),          # ragged last tile, qkv-backward shape of stage 0
    (128 * 300 + 37, 96, 384, True, True),  # several tiles per CTA, fc1-backward shape, DropPath-scaled second output
    (777, 192, 576, False, True),          # two boxes per warp, no residual-path gradient
    (4096, 192, 768, True, False),
])
def test_gemm_layernorm_backward_epilogue(M, C, K, res, scaled):
    """EPI_LNBWD (gemm.cuh): dX GEMM whose epilogue is the LayerNorm backward of tulip.py:338,348 (norm1 / norm2) -- against the
    same formulas in fp32 torch on the same bf16 operands.  dx is rounded once (1e-3 against the bf16-rounded fp32 result, the
    cancellation in g - mean(g) - xhat mean(g xhat) included); d(gamma), d(beta) are fp32 sums of fp32 accumulators (1e-4)."""
    from tulip_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(M + C + K)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=gen) * scale)
    A, W = bf(rnd(M, K).cuda()), bf(rnd(C, K, scale=K ** -0.5).cuda())
    x = bf((rnd(M, C, scale=1.5) + 0.3).cuda())
    dres = bf(rnd(M, C).cuda()) if res else None
    gamma = (1.0 + 0.2 * rnd(C)).cuda().contiguous()
    eps = 1e-5
    xf = x.float()
    mean = xf.mean(-1)
    rstd = torch.rsqrt(xf.var(-1, unbiased=False) + eps)
    stats = torch.stack([mean, rstd], dim=-1).contiguous()
    rps = 64
    scale = (torch.rand((M + rps - 1) // rps, generator=gen) + 0.5).cuda().contiguous() if scaled else None

    dy = A.float() @ W.float().t()
    xh = (xf - mean[:, None]) * rstd[:, None]
    g = dy * gamma
    want = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if res:
        want = want + dres.float()
    want_dw, want_db = (dy * xh).sum(0), dy.sum(0)

    dx = torch.full((M, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    dxs = torch.full((M, C), float("nan"), dtype=torch.bfloat16, device="cuda") if scaled else None
    dw = torch.zeros(C, dtype=torch.float32, device="cuda")
    db = torch.zeros(C, dtype=torch.float32, device="cuda")
    kw = dict(A=A, lda=K, K1=K, B=W, ldb=K, M=M, N=C, K=K, out=dx, ldo=C, aux=x, ldaux=C, ln_w=gamma, ln_stats=stats, ln_dw=dw, ln_db=db)
    if res:
        kw.update(aux2=dres, ldaux2=C)
    if scaled:
        kw.update(out2=dxs, ldo2=C, row_scale=scale, rows_per_sample=rps)
    ops.gemm_nt_ex(ops.EPI_LNBWD, **kw)
    torch.cuda.synchronize()
    assert torch.isfinite(dx.float()).all()
    assert rel_l2(dx.float(), bf16r(want)) <= 1e-3
    assert rel_l2(dw, want_dw) <= 1e-4 and rel_l2(db, want_db) <= 1e-4
    if scaled:
        rows = torch.arange(M, device="cuda") // rps
        assert torch.equal(dxs, bf(scale[rows][:, None] * dx.float()))


@pytest.mark.parametrize("M,C,K,resid", [
    (1000, 96, 192, False),                 # skip-Linear shape of stage 0 (ragged last tile)
    (128 * 200 + 5, 96, 96, True),          # proj + residual, several tiles per CTA
    (777, 192, 768, True),                  # fc2 + residual of stage 1: two boxes per warp
    (4096, 192, 384, False),                # PatchMerging reduction feeding stage 1
])
def test_gemm_layernorm_forward_epilogue(M, C, K, resid):
    """EPI_STORE_LN / EPI_RESID_LN (gemm.cuh): the Linear (+ residual + DropPath scale) and the LayerNorm that reads its output
    (tulip.py:338,348: norm1 / norm2 of the next half-block) in one launch.  out must equal the plain epilogue's result bit for
    bit; ln_y / statistics are checked against fp32 torch LayerNorm of that (bf16) output: one rounding, 1e-3."""
    from tulip_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(M + C + K)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=gen) * scale)
    A, W = bf(rnd(M, K).cuda()), bf(rnd(C, K, scale=K ** -0.5).cuda())
    bias = rnd(C, scale=0.1).cuda().contiguous()
    gamma, beta = (1.0 + 0.2 * rnd(C)).cuda().contiguous(), (0.1 * rnd(C)).cuda().contiguous()
    aux = bf((rnd(M, C, scale=2.0) + 0.5).cuda()) if resid else None
    rps = 50
    scale = (torch.rand((M + rps - 1) // rps, generator=gen) + 0.5).cuda().contiguous() if resid else None
    eps = 1e-5
    base = dict(A=A, lda=K, K1=K, B=W, ldb=K, M=M, N=C, K=K, bias=bias, ldo=C)
    if resid:
        base.update(aux=aux, ldaux=C, row_scale=scale, rows_per_sample=rps)
    ref = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_RESID if resid else ops.EPI_STORE, out=ref, **base)
    out = torch.full((M, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    y = torch.full((M, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats = torch.full((M, 2), float("nan"), dtype=torch.float32, device="cuda")
    ops.gemm_nt_ex(ops.EPI_RESID_LN if resid else ops.EPI_STORE_LN, out=out, ln_w=gamma, ln_b=beta, ln_y=y, ln_ystats=stats, ln_eps=eps,
                   **base)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    xf = out.float()
    want = torch.nn.functional.layer_norm(xf, (C,), gamma, beta, eps)
    assert rel_l2(y.float(), bf16r(want)) <= 1e-3
    mean, var = xf.mean(-1), xf.var(-1, unbiased=False)
    assert torch.allclose(stats[:, 0], mean, rtol=1e-5, atol=1e-5)
    assert torch.allclose(stats[:, 1], torch.rsqrt(var + eps), rtol=1e-4, atol=0)


def test_final_patch_expanding_head_fwd_bwd():
    """FinalPatchExpanding + decoder_pred + L1 (tulip.py:144-159, 727-731, 692-693) as the head epilogues 6 / 7 with hd_ln:
    Linear(E -> r^2 E, no bias) -> '(P1 P2 C)' rearrange -> LayerNorm(E) per output pixel -> 1x1 conv E -> 1.
    Checked against the same math in fp32 torch from the same bf16 operands (tolerances as for the PixelShuffle head: pred 1e-5,
    fp32 gradient accumulators 2e-3, d(acc) stored as bf16 3e-3)."""
    from tulip_b200 import ops
    B, H, W, E, r = 2, 4, 32, 96, 4
    T, N = B * H * W, E * r * r
    g = torch.Generator().manual_seed(5)
    rb = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).bfloat16().float()
    xn, we = rb(T, E), rb(N, E, sc=0.1)
    gam, bet, wd = (1 + 0.1 * torch.randn(E, generator=g)), 0.1 * torch.randn(E, generator=g), 0.2 * torch.randn(E, generator=g)
    target = rb(B, 1, H * r, W * r, sc=0.5)
    cu = lambda t: t.cuda().contiguous()
    xb, wb = cu(xn).bfloat16(), cu(we).bfloat16()
    pred = torch.zeros((B, 1, H * r, W * r), dtype=torch.float32, device="cuda")
    stats = torch.zeros((B * H * r * W * r, 2), dtype=torch.float32, device="cuda")
    gam_c, bet_c, wd_c, tgt_c = cu(gam), cu(bet), cu(wd), cu(target)
    ops.gemm_nt_ex(ops.EPI_HEAD, A=xb, lda=E, K1=E, B=wb, ldb=E, M=T, N=N, K=E, wd=wd_c, pred=pred, hd_H=H, hd_W=W, hd_r=r, hd_E=E,
                   hd_ln=1, ln_w=gam_c, ln_b=bet_c, ln_eps=1e-6, ln_ystats=stats)
    # fp32 torch on the same operands
    xf, wf = xn.clone().requires_grad_(True), we.clone().requires_grad_(True)
    gf, bf_, wdf = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True), wd.clone().requires_grad_(True)
    y = torch.nn.functional.linear(xf, wf).view(B, H, W, r, r, E).permute(0, 1, 3, 2, 4, 5).reshape(B, H * r, W * r, E)
    yn = torch.nn.functional.layer_norm(y, (E,), gf, bf_, 1e-6)
    pred_t = (yn * wdf).sum(-1).unsqueeze(1)
    assert rel_l2(pred, pred_t) <= 1e-5
    mean_t = y.mean(-1).reshape(-1)
    assert rel_l2(stats[:, 0], mean_t) <= 1e-5
    assert rel_l2(stats[:, 1], (y.var(-1, unbiased=False) + 1e-6).rsqrt().reshape(-1)) <= 1e-5
    dpred = torch.sign(pred.cpu() - target) / pred.numel()
    pred_t.backward(dpred)
    dh = torch.empty((T, N), dtype=torch.bfloat16, device="cuda")
    dg, db, dwd = (torch.zeros(E, dtype=torch.float32, device="cuda") for _ in range(3))
    gscale = torch.ones(1, dtype=torch.float32, device="cuda")
    ops.gemm_nt_ex(ops.EPI_HEAD_BWD, A=xb, lda=E, K1=E, B=wb, ldb=E, M=T, N=N, K=E, wd=wd_c, pred=pred, target=tgt_c, gscale=gscale,
                   dwd=dwd, out=dh, ldo=N, hd_H=H, hd_W=W, hd_r=r, hd_E=E, hd_ln=1, ln_w=gam_c, ln_b=bet_c, ln_eps=1e-6, ln_stats=stats,
                   ln_dw=dg, ln_db=db)
    assert rel_l2(dg, gf.grad) <= 2e-3 and rel_l2(db, bf_.grad) <= 2e-3 and rel_l2(dwd, wdf.grad) <= 2e-3
    dWe = torch.zeros((N, E), dtype=torch.float32, device="cuda")
    ops.gemm_tn_ex(dY=dh, ldy=N, X=xb, ldx=E, K1=E, M=T, N=N, K=E, dW=dWe, lddw=E)
    dxn = torch.empty((T, E), dtype=torch.bfloat16, device="cuda")
    wtb = cu(we.t()).bfloat16()
    ops.gemm_nt_ex(ops.EPI_STORE, A=dh, lda=N, K1=N, B=wtb, ldb=N, M=T, N=E, K=N, out=dxn, ldo=E)
    assert rel_l2(dWe, wf.grad) <= 3e-3
    assert rel_l2(dxn.float(), xf.grad) <= 3e-3


@pytest.mark.parametrize("M,C", [(4096, 96), (128 * 160, 96), (128 * 152 + 77, 96), (128 * 150, 192), (9000, 384)],
                         ids=["c96", "c96_full_waves", "c96_ragged", "c192", "c384"])
def test_gemm_nt_dgelu2_recompute(M, C):
    """EPI_DGELU2 (Mlp backward through GELU, tulip.py:195-199 differentiated): out = (dY . W2^T-packed) * gelu'(xn . W1^T + b1),
    both products in one launch (two TMEM accumulators per tile), incl. a ragged last row tile and more tiles than CTAs.
    One bf16 rounding after fp32 accumulation: rel-L2 <= 1e-3 against the fp32 torch reference rounded to bf16."""
    from tulip_b200 import ops
    N = 4 * C
    g = torch.Generator(device="cuda").manual_seed(M + C)
    dy = bf(torch.randn(M, C, device="cuda", generator=g))
    xn = bf(torch.randn(M, C, device="cuda", generator=g))
    w2t = bf(torch.randn(N, C, device="cuda", generator=g) / C ** 0.5)      # fc2.weight^T: [4C, C]
    w1 = bf(torch.randn(N, C, device="cuda", generator=g) / C ** 0.5)       # fc1.weight:   [4C, C]
    b1 = (0.2 * torch.randn(N, device="cuda", generator=g)).contiguous()
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.gemm_nt_ex(ops.EPI_DGELU2, A=dy, lda=C, K1=C, B=w2t, ldb=C, A2=xn, lda2=C, B2=w1, ldb2=C, K2=C, M=M, N=N, K=C,
                   bias=b1, out=out, ldo=N)
    pre = xn.float() @ w1.float().t() + b1
    gp = 0.5 * (1 + torch.erf(pre / 2 ** 0.5)) + pre * torch.exp(-0.5 * pre * pre) / (2 * torch.pi) ** 0.5
    ref = bf16r((dy.float() @ w2t.float().t()) * gp)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) <= 1e-3
