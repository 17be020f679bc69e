"""Thin torch-tensor wrappers over the per-kernel C ABI (include/tulip_b200.h).

Every function takes CUDA tensors, enqueues on torch's current stream and returns new tensors.
Activations are bf16; parameters, statistics and gradients fp32.  No fallbacks: a missing library or
a CPU tensor raises.
"""
from __future__ import annotations

import torch

import ctypes as _C

from ._lib import GemmDesc, GemmTNDesc, check, current_stream, load_library, ptr

GEMM_AUTO, GEMM_TC05 = 0, 2      # one implementation (tcgen05); 2 is kept as an alias of 0
EPI_STORE, EPI_GELU, EPI_RESID, EPI_PIXSHUF, EPI_SPLIT2, EPI_DGELU, EPI_HEAD, EPI_HEAD_BWD, EPI_ROWSCALE, EPI_DGELU2, EPI_LNBWD, EPI_STORE_LN, EPI_RESID_LN = range(13)
A_PLAIN, A_UNSHUFFLE = 0, 1


def gemm_nt_ex(epilogue, **fields):
    """tulip_gemm_nt_ex: every operand mode / fused epilogue of the NT contraction.  Tensor-valued fields are passed as tensors
    (the caller keeps them alive and supplies the leading dimensions)."""
    d = GemmDesc()
    for k, v in fields.items():
        setattr(d, k, ptr(v) if isinstance(v, torch.Tensor) else v)
    check(load_library().tulip_gemm_nt_ex(_C.byref(d), epilogue, current_stream()), "tulip_gemm_nt_ex")


def gemm_tn_ex(**fields):
    d = GemmTNDesc()
    for k, v in fields.items():
        setattr(d, k, ptr(v) if isinstance(v, torch.Tensor) else v)
    check(load_library().tulip_gemm_tn_ex(_C.byref(d), current_stream()), "tulip_gemm_tn_ex")


def gemm_tn_group(problems):
    """tulip_gemm_tn_group: up to 4 weight gradients dW_p += dY_p^T X_p (db_p += colsum dY_p) in one persistent launch.
    problems: list of dicts with the tulip_gemm_tn_desc fields (tensors for the pointers)."""
    arr = (GemmTNDesc * len(problems))()
    for d, fields in zip(arr, problems):
        for k, v in fields.items():
            setattr(d, k, ptr(v) if isinstance(v, torch.Tensor) else v)
    check(load_library().tulip_gemm_tn_group(arr, len(problems), current_stream()), "tulip_gemm_tn_group")


def gemm_tn_group_plan(shapes, sms=148):
    """Host-side work-item cut of a grouped launch: shapes = [(M, N, K), ...] -> (64-token blocks per item, items)."""
    n = len(shapes)
    I = _C.c_int * n
    per, items = (_C.c_int * 4)(), _C.c_int(0)
    check(load_library().tulip_gemm_tn_group_plan(I(*[s[0] for s in shapes]), I(*[s[1] for s in shapes]), I(*[s[2] for s in shapes]),
                                                  n, sms, per, _C.byref(items)), "tulip_gemm_tn_group_plan")
    return list(per)[:n], items.value


def permute_rows_for_shuffle(w, R2, Cc):
    """Row order the PixelShuffle-feeding GEMMs use: destination row n' = ij*Cc + c holds source row c*R2 + ij."""
    return w.reshape(Cc, R2, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tulip_b200.ops: CUDA tensors only (no CPU fallback)")


def _bf16(t):
    return t.to(torch.bfloat16).contiguous()


def _f32(t):
    return None if t is None else t.to(torch.float32).contiguous()


def linear(x, w, bias=None, epilogue=EPI_STORE, aux=None, row_scale=None, rows_per_sample=1, impl=GEMM_AUTO, save_pre=True):
    """x [M,K] bf16, w [N,K] bf16 -> [M,N] bf16; EPI_GELU returns (gelu(pre), pre) -- pre is None with save_pre=False
    (the configuration the network executor uses: the backward pass recomputes the pre-activation)."""
    _cuda(x, w)
    x, w = _bf16(x), _bf16(w)
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    out2 = torch.empty_like(out) if (epilogue == EPI_GELU and save_pre) else None
    aux = None if aux is None else _bf16(aux)
    bias, row_scale = _f32(bias), _f32(row_scale)
    check(load_library().tulip_gemm_nt(ptr(x), ptr(w), ptr(bias), ptr(out), ptr(out2), ptr(aux), ptr(row_scale), rows_per_sample,
                                       M, N, K, epilogue, impl, current_stream()), "tulip_gemm_nt")
    return (out, out2) if epilogue == EPI_GELU else out


def linear_wgrad(dy, x, want_bias=True, impl=GEMM_AUTO):
    """dW [N,K] = dy[M,N]^T x[M,K] (fp32), db [N] = colsum(dy)."""
    _cuda(dy, x)
    dy, x = _bf16(dy), _bf16(x)
    M, N = dy.shape
    K = x.shape[1]
    dW = torch.zeros((N, K), dtype=torch.float32, device=x.device)
    db = torch.zeros(N, dtype=torch.float32, device=x.device) if want_bias else None
    check(load_library().tulip_gemm_tn(ptr(dy), ptr(x), ptr(dW), ptr(db), M, N, K, impl, current_stream()), "tulip_gemm_tn")
    return dW, db


def wmsa_block(x, ln_w, ln_b, wqkv, bqkv, wproj, bproj, bias_table, B, H, W, heads, window=(2, 8), shift=(0, 0), masked=False,
               bias_window=(2, 8), row_scale=None, eps=1e-6):
    """Fused attention half-block: x [B*H*W, C] bf16 -> x + row_scale[b] * proj(attn(qkv(LayerNorm(x)))) (tulip.py:338-346)."""
    _cuda(x, wqkv, wproj)
    x, wqkv, wproj = _bf16(x), _bf16(wqkv), _bf16(wproj)
    ln_w, ln_b, bqkv, bproj, bias_table, row_scale = (_f32(t) for t in (ln_w, ln_b, bqkv, bproj, bias_table, row_scale))
    C = x.shape[1]
    y = torch.empty_like(x)
    check(load_library().tulip_wmsa_block_fwd(ptr(x), ptr(y), ptr(ln_w), ptr(ln_b), ptr(wqkv), ptr(bqkv), ptr(wproj), ptr(bproj),
                                              ptr(bias_table), ptr(row_scale), B, H, W, C, heads, window[0], window[1], shift[0],
                                              shift[1], int(masked), bias_window[0], bias_window[1], float(eps), current_stream()),
          "tulip_wmsa_block_fwd")
    return y


def mlp_block(x, ln_w, ln_b, w1, b1, w2, b2, row_scale=None, rows_per_sample=1, save=False, eps=1e-6):
    """Fused MLP half-block: x [T, C] bf16 -> x + row_scale[sample] * fc2(gelu(fc1(LayerNorm(x)))) (tulip.py:347-352).
    save=True also returns (xn, stats, hact), what the backward pass consumes."""
    _cuda(x, w1, w2)
    x, w1, w2 = _bf16(x), _bf16(w1), _bf16(w2)
    ln_w, ln_b, b1, b2, row_scale = (_f32(t) for t in (ln_w, ln_b, b1, b2, row_scale))
    T, C = x.shape
    y = torch.empty_like(x)
    xn = torch.empty_like(x) if save else None
    stats = torch.empty((T, 2), dtype=torch.float32, device=x.device) if save else None
    hact = torch.empty((T, 4 * C), dtype=torch.bfloat16, device=x.device) if save else None
    check(load_library().tulip_mlp_block_fwd(ptr(x), ptr(y), ptr(ln_w), ptr(ln_b), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(row_scale),
                                             rows_per_sample, ptr(xn), ptr(stats), ptr(hact), T, C, float(eps), current_stream()),
          "tulip_mlp_block_fwd")
    return (y, xn, stats, hact) if save else y


def window_attention(qkv, bias_table, B, H, W, heads, window=(2, 8), shift=(0, 0), masked=False, bias_window=(2, 8)):
    """qkv [B*H*W, 3C] bf16 (natural token order) -> [B*H*W, C] bf16."""
    _cuda(qkv, bias_table)
    qkv, bias_table = _bf16(qkv), _f32(bias_table)
    C = qkv.shape[1] // 3
    out = torch.empty((qkv.shape[0], C), dtype=torch.bfloat16, device=qkv.device)
    check(load_library().tulip_window_attention_fwd(ptr(qkv), ptr(bias_table), ptr(out), B, H, W, C, heads, window[0], window[1],
                                                    shift[0], shift[1], int(masked), bias_window[0], bias_window[1],
                                                    current_stream()), "tulip_window_attention_fwd")
    return out


def window_attention_bwd(qkv, bias_table, dout, B, H, W, heads, window=(2, 8), shift=(0, 0), masked=False, bias_window=(2, 8)):
    _cuda(qkv, bias_table, dout)
    qkv, bias_table, dout = _bf16(qkv), _f32(bias_table), _bf16(dout)
    C = qkv.shape[1] // 3
    dqkv = torch.empty_like(qkv)
    dtable = torch.zeros_like(bias_table)
    check(load_library().tulip_window_attention_bwd(ptr(qkv), ptr(bias_table), ptr(dout), ptr(dqkv), ptr(dtable), B, H, W, C, heads,
                                                    window[0], window[1], shift[0], shift[1], int(masked), bias_window[0],
                                                    bias_window[1], current_stream()), "tulip_window_attention_bwd")
    return dqkv, dtable


def layernorm(x, w, b, eps=1e-6, merge=None):
    """x [rows, C] bf16 -> (y bf16, stats [rows,2] fp32).  merge=(B,H,W): x is [B,H,W,Cs] and rows are the
    2x2-gathered 4*Cs vectors of PatchMerging."""
    _cuda(x, w, b)
    x, w, b = _bf16(x), _f32(w), _f32(b)
    if merge is None:
        rows, C = x.shape
        g, H2, W2 = 0, 0, 0
    else:
        B, H, W = merge
        C = 4 * x.shape[-1]
        rows, g, H2, W2 = B * (H // 2) * (W // 2), 1, H // 2, W // 2
    y = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
    check(load_library().tulip_layernorm_fwd(ptr(x), ptr(w), ptr(b), ptr(y), ptr(stats), rows, C, eps, g, H2, W2, current_stream()),
          "tulip_layernorm_fwd")
    return y, stats


def layernorm_bwd(x, w, stats, dy, dres=None, merge=None):
    _cuda(x, w, stats, dy)
    x, w, dy = _bf16(x), _f32(w), _bf16(dy)
    rows, C = dy.shape
    g, H2, W2 = (0, 0, 0) if merge is None else (1, merge[1] // 2, merge[2] // 2)
    dx = torch.empty_like(x)
    dw = torch.zeros(C, dtype=torch.float32, device=x.device)
    db = torch.zeros(C, dtype=torch.float32, device=x.device)
    dres = None if dres is None else _bf16(dres)
    check(load_library().tulip_layernorm_bwd(ptr(x), ptr(w), ptr(stats), ptr(dy), ptr(dres), ptr(dx), ptr(dw), ptr(db), rows, C,
                                             g, H2, W2, current_stream()), "tulip_layernorm_bwd")
    return dx, dw, db


def patch_embed(x, w, b, ln_w, ln_b, eps=1e-6):
    """x [B,1,H,W] fp32 -> [B, H/ph, W/4, E] bf16."""
    _cuda(x, w)
    x, w, b, ln_w, ln_b = _f32(x), _f32(w), _f32(b), _f32(ln_w), _f32(ln_b)
    B, _, H, W = x.shape
    E, ph = w.shape[0], w.shape[2]
    y = torch.empty((B, H // ph, W // 4, E), dtype=torch.bfloat16, device=x.device)
    check(load_library().tulip_patch_embed_fwd(ptr(x), ptr(w), ptr(b), ptr(ln_w), ptr(ln_b), ptr(y), B, H, W, ph, E, eps,
                                               current_stream()), "tulip_patch_embed_fwd")
    return y


def patch_embed_bwd(x, w, b, ln_w, dy, eps=1e-6):
    _cuda(x, w, dy)
    x, w, b, ln_w, dy = _f32(x), _f32(w), _f32(b), _f32(ln_w), _bf16(dy)
    B, _, H, W = x.shape
    E, ph = w.shape[0], w.shape[2]
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    dlw, dlb = torch.zeros_like(ln_w), torch.zeros_like(ln_w)
    check(load_library().tulip_patch_embed_bwd(ptr(x), ptr(w), ptr(b), ptr(ln_w), ptr(dy), ptr(dw), ptr(db), ptr(dlw), ptr(dlb),
                                               B, H, W, ph, E, eps, current_stream()), "tulip_patch_embed_bwd")
    return dw, db, dlw, dlb


def l1_loss(pred, target, log_transform=True):
    _cuda(pred, target)
    pred, target = _f32(pred), _f32(target)
    scratch = torch.zeros(2, dtype=torch.float32, device=pred.device)
    out = torch.zeros(2, dtype=torch.float32, device=pred.device)
    check(load_library().tulip_l1_loss(ptr(pred), ptr(target), pred.numel(), int(log_transform), ptr(scratch), ptr(out),
                                       current_stream()), "tulip_l1_loss")
    return out[0], out[1]


def eval_postprocess(pred, x_lo, target, log_transform=True, clip_lo=2.0 / 80.0, keep_low_res=True):
    """Fused evaluation post-processing (reference engine_upsampling.py:174-244): returns (range image [B,1,H,W] fp32 with the
    sensor rows restored, losses [B,2] = per-frame {pixel loss, loss on the sensor rows})."""
    _cuda(pred, x_lo, target)
    pred, x_lo, target = _f32(pred), _f32(x_lo), _f32(target)
    B, _, H, W = pred.shape
    if keep_low_res and x_lo.shape[-1] != W:
        raise ValueError("eval_postprocess: keep_low_res needs input and target of equal width")
    out = torch.empty_like(pred)
    losses = torch.empty((B, 2), dtype=torch.float32, device=pred.device)
    scratch = torch.empty(2 * B, dtype=torch.float32, device=pred.device)
    check(load_library().tulip_eval_postprocess(ptr(pred), ptr(x_lo), ptr(target), ptr(out), ptr(losses), ptr(scratch), B, H, W,
                                                x_lo.shape[2], int(log_transform), float(clip_lo), int(keep_low_res),
                                                current_stream()), "tulip_eval_postprocess")
    return out, losses


def mc_dropout_aggregate(preds, noise_threshold=0.03, return_std=False):
    """preds [N, 1, H, W] fp32 (stacked mc_drop passes of one frame) -> [1, 1, H, W]: mean over the passes with the pixels whose
    unbiased std exceeds noise_threshold * mean zeroed (engine_upsampling.py:423-427)."""
    _cuda(preds)
    preds = _f32(preds)
    n = preds.shape[0]
    out = torch.empty((1, *preds.shape[1:]), dtype=torch.float32, device=preds.device)
    std = torch.empty_like(out) if return_std else None
    check(load_library().tulip_mc_dropout_aggregate(ptr(preds), ptr(out), ptr(std), n, out.numel(), float(noise_threshold),
                                                    current_stream()), "tulip_mc_dropout_aggregate")
    return (out, std) if return_std else out


def window_partition(x, window=(2, 8), shift=(0, 0)):
    """(B,H,W,C) bf16 -> ((B Nh Nw), Mh, Mw, C): torch.roll(x, (-sh,-sw)) then the reference's window_partition."""
    _cuda(x)
    x = _bf16(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B * (H // window[0]) * (W // window[1]), window[0], window[1], Cc), dtype=torch.bfloat16, device=x.device)
    check(load_library().tulip_window_partition(ptr(x), ptr(out), B, H, W, Cc, window[0], window[1], shift[0], shift[1],
                                                current_stream()), "tulip_window_partition")
    return out


def window_reverse(xw, B, H, W, window=(2, 8), shift=(0, 0)):
    _cuda(xw)
    xw = _bf16(xw)
    Cc = xw.shape[-1]
    out = torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=xw.device)
    check(load_library().tulip_window_reverse(ptr(xw), ptr(out), B, H, W, Cc, window[0], window[1], shift[0], shift[1],
                                              current_stream()), "tulip_window_reverse")
    return out


def shift_mask(H, W, window=(2, 8), shift=(1, 4), device="cuda"):
    L = window[0] * window[1]
    out = torch.empty(((H // window[0]) * (W // window[1]), L, L), dtype=torch.float32, device=device)
    check(load_library().tulip_shift_mask(ptr(out), H, W, window[0], window[1], shift[0], shift[1], current_stream()),
          "tulip_shift_mask")
    return out


def rel_bias_gather(table, window=(2, 8)):
    _cuda(table)
    table = _f32(table)
    L = window[0] * window[1]
    out = torch.empty((table.shape[1], L, L), dtype=torch.float32, device=table.device)
    check(load_library().tulip_rel_bias_gather(ptr(table), ptr(out), table.shape[1], window[0], window[1], current_stream()),
          "tulip_rel_bias_gather")
    return out


def merge_gather(x):
    _cuda(x)
    x = _bf16(x)
    B, H, W, Cc = x.shape
    out = torch.empty((B, H // 2, W // 2, 4 * Cc), dtype=torch.bfloat16, device=x.device)
    check(load_library().tulip_merge_gather(ptr(x), ptr(out), B, H, W, Cc, current_stream()), "tulip_merge_gather")
    return out


def pixel_shuffle(x, r):
    """NHWC PixelShuffle: (B,H,W,Cout*r*r) -> (B,H*r,W*r,Cout)."""
    _cuda(x)
    x = _bf16(x)
    B, H, W, Crr = x.shape
    out = torch.empty((B, H * r, W * r, Crr // (r * r)), dtype=torch.bfloat16, device=x.device)
    check(load_library().tulip_pixel_shuffle(ptr(x), ptr(out), B, H, W, Crr // (r * r), r, current_stream()), "tulip_pixel_shuffle")
    return out
