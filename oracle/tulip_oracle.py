"""Functional fp32 torch restatement of the TULIP Swin U-Net forward pass.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- the floating-point oracle the CUDA
path is compared with.  Parameters come in as a plain ``dict`` keyed like the
reference ``state_dict``; there are no nn.Modules and no hidden state (the one
stateful quirk of the reference, the permanent backup-window switch, is carried
in an explicit ``state`` dict).  Gradients are obtained with torch.autograd on
these functions (the reference defines no backward either: SURVEY.md App. G).

Every function cites the reference lines it restates (relative to
/root/reference).  Pinned against the unmodified reference by
oracle/make_golden.py -> tests/golden/*.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .params import Cfg

# ----------------------------------------------------------------------------- index helpers


def window_partition(x: torch.Tensor, win) -> torch.Tensor:
    """tulip/model/tulip.py:248-252."""
    B, H, W, C = x.shape
    Mh, Mw = win
    return x.view(B, H // Mh, Mh, W // Mw, Mw, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, Mh * Mw, C)


def window_reverse(xw: torch.Tensor, win, B: int, H: int, W: int) -> torch.Tensor:
    """tulip/model/tulip.py:320."""
    Mh, Mw = win
    C = xw.shape[-1]
    return xw.view(B, H // Mh, W // Mw, Mh, Mw, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


def shift_mask(H: int, W: int, win, shift, device=None) -> torch.Tensor:
    """tulip/model/tulip.py:254-280 (closed form; equality with the slice-fill
    procedure is tested in tests/test_oracle_index_ops.py)."""
    from .index_ops import shift_mask_closed_form
    return torch.from_numpy(shift_mask_closed_form(H, W, tuple(win), tuple(shift))).to(device)


def relative_position_bias(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """(nbias, heads) table + (L, L) index -> (heads, L, L). tulip.py:304-307."""
    L = index.shape[0]
    return table[index.reshape(-1)].view(L, L, -1).permute(2, 0, 1)


# ----------------------------------------------------------------------------- layers


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def patch_embed(x: torch.Tensor, p: dict, cfg: Cfg) -> torch.Tensor:
    """Circular pad W by (2,2) -> Conv2d(k=(ph,8), s=(ph,pw)) -> NHWC -> LayerNorm.
    tulip/model/tulip.py:59-73 (the zero `padding` at :50-56 is a no-op for all
    shapes that divide by the patch size, which is asserted here)."""
    ph, pw = cfg.patch_size
    assert x.shape[2] % ph == 0 and x.shape[3] % pw == 0
    xp = torch.cat([x[..., -2:], x, x[..., :2]], dim=-1)
    y = F.conv2d(xp, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=(ph, pw))
    y = y.permute(0, 2, 3, 1)
    return layer_norm(y, p["patch_embed.norm.weight"], p["patch_embed.norm.bias"], cfg.ln_eps)


def effective_window(H: int, win, shift_flag: bool, state: dict | None, key: str):
    """Window / shift actually used by one WindowAttention call, including the
    permanent backup switch of tulip.py:216-222, 284-287."""
    win = tuple(win)
    L = win[0] * win[1]
    cur = state.get(key, win) if state is not None else win
    if H < cur[0]:
        cur = (1, L)
        if state is not None:
            state[key] = cur
    if not shift_flag:
        return cur, (0, 0)
    if cur == win:
        return cur, (win[0] // 2, win[1] // 2)
    return cur, (0, L // 2)


def window_attention(x, p, prefix, heads, win, shift_flag, state=None):
    """x: (B,H,W,C) already LayerNorm'ed -> (B,H,W,C). tulip/model/tulip.py:282-324."""
    B, H, W, C = x.shape
    cur, (sh, sw) = effective_window(H, win, shift_flag, state, prefix)
    hd = C // heads
    scale = (C // heads) ** -0.5
    if shift_flag:
        x = torch.roll(x, shifts=(-sh, -sw), dims=(1, 2))
        mask = shift_mask(H, W, cur, (sh, sw), x.device)
    xw = window_partition(x, cur)                                     # (Bn, L, C)
    Bn, L, _ = xw.shape
    qkv = F.linear(xw, p[f"{prefix}.qkv.weight"], p[f"{prefix}.qkv.bias"])
    qkv = qkv.view(Bn, L, 3, heads, hd).permute(2, 0, 3, 1, 4)        # (3, Bn, h, L, hd): f = t*C + head*hd + d
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    attn = attn + relative_position_bias(p[f"{prefix}.relative_position_bias_table"],
                                         p[f"{prefix}.relative_position_index"]).unsqueeze(0)
    if shift_flag:
        nW = mask.shape[0]
        attn = (attn.view(Bn // nW, nW, heads, L, L) + mask[None, :, None]).view(Bn, heads, L, L)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(Bn, L, C)              # head-major channel concat (:317)
    o = F.linear(o, p[f"{prefix}.proj.weight"], p[f"{prefix}.proj.bias"])
    o = window_reverse(o, cur, B, H, W)
    if shift_flag:
        o = torch.roll(o, shifts=(sh, sw), dims=(1, 2))
    return o


def mlp(x, p, prefix):
    """fc1 -> exact-erf GELU -> fc2. tulip/model/tulip.py:194-200."""
    h = F.linear(x, p[f"{prefix}.fc1.weight"], p[f"{prefix}.fc1.bias"])
    h = F.gelu(h)
    return F.linear(h, p[f"{prefix}.fc2.weight"], p[f"{prefix}.fc2.bias"])


def drop_path_scale(x, scale):
    """DropPath with an explicit per-sample scale (0 or 1/keep); tulip.py:16-30."""
    if scale is None:
        return x
    return x * scale.view(-1, *([1] * (x.ndim - 1))).to(x.dtype)


def swin_block(x, p, prefix, heads, win, shift_flag, cfg, state=None, dp=(None, None)):
    """Pre-norm residual block. tulip/model/tulip.py:338-352."""
    y = layer_norm(x, p[f"{prefix}.norm1.weight"], p[f"{prefix}.norm1.bias"], cfg.ln_eps)
    y = window_attention(y, p, f"{prefix}.attn", heads, win, shift_flag, state)
    x = x + drop_path_scale(y, dp[0])
    y = layer_norm(x, p[f"{prefix}.norm2.weight"], p[f"{prefix}.norm2.bias"], cfg.ln_eps)
    y = mlp(y, p, f"{prefix}.mlp")
    return x + drop_path_scale(y, dp[1])


def patch_merging(x, p, prefix, cfg):
    """2x2 gather [(0,0),(1,0),(0,1),(1,1)] -> LN(4C) -> Linear(4C->2C, no bias). tulip.py:92-106."""
    assert x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], dim=-1)
    x = layer_norm(x, p[f"{prefix}.norm.weight"], p[f"{prefix}.norm.bias"], cfg.ln_eps)
    return F.linear(x, p[f"{prefix}.reduction.weight"])


def patch_unmerging(x, p, prefix):
    """1x1 conv C->2C (+bias) then PixelShuffle(2), all in NHWC. tulip.py:117-123.
    out[b,2h+i,2w+j,c] = (x[b,h,w,:] . W[4c+2i+j,:]) + bias[4c+2i+j]."""
    B, H, W, C = x.shape
    w = p[f"{prefix}.expand.weight"].view(2 * C, C)
    y = F.linear(x, w, p[f"{prefix}.expand.bias"])                    # (B,H,W,2C)
    y = y.view(B, H, W, C // 2, 2, 2).permute(0, 1, 4, 2, 5, 3)
    return y.reshape(B, 2 * H, 2 * W, C // 2)


def patch_expanding(x, p, prefix, cfg):
    """Linear(C -> 2C, no bias) -> 'B H W (P1 P2 C) -> B (H P1) (W P2) C' with P1 = P2 = 2 -> LayerNorm(C/2).
    tulip.py:126-141.  out[b,2h+i,2w+j,:] = LN(x[b,h,w,:] . W[(2i+j)*C/2 : (2i+j+1)*C/2, :]^T)."""
    B, H, W, C = x.shape
    y = F.linear(x, p[f"{prefix}.expand.weight"])                      # (B,H,W,2C), columns (P1 P2 C)
    y = y.view(B, H, W, 2, 2, C // 2).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C // 2)
    return layer_norm(y, p[f"{prefix}.norm.weight"], p[f"{prefix}.norm.bias"], cfg.ln_eps)


def head_expanding(x, p, cfg):
    """norm_up -> FinalPatchExpanding (Linear(E -> r^2 E, no bias) -> '(P1 P2 C)' rearrange -> LayerNorm(E)) -> decoder_pred.
    tulip.py:720, 727-731, 144-159."""
    B, H, W, E = x.shape
    r = cfg.upscale_factor
    x = layer_norm(x, p["norm_up.weight"], p["norm_up.bias"], cfg.ln_eps)
    y = F.linear(x, p["final_patch_expanding.expand.weight"])
    y = y.view(B, H, W, r, r, E).permute(0, 1, 3, 2, 4, 5).reshape(B, H * r, W * r, E)
    y = layer_norm(y, p["final_patch_expanding.norm.weight"], p["final_patch_expanding.norm.bias"], cfg.ln_eps)
    y = F.linear(y, p["decoder_pred.weight"].view(cfg.in_chans, E))
    return y.permute(0, 3, 1, 2)


def head(x, p, cfg):
    """norm_up -> 1x1 conv E->E*r^2 (+bias) -> LeakyReLU(0.01) -> PixelShuffle(r) -> 1x1 conv E->in_chans.
    tulip/model/tulip.py:720-731, 161-178, 574.  x: (B,H,W,E) -> (B,in_chans,H*r,W*r)."""
    B, H, W, E = x.shape
    r = cfg.upscale_factor
    x = layer_norm(x, p["norm_up.weight"], p["norm_up.bias"], cfg.ln_eps)
    y = F.linear(x, p["ps_head.conv_expand.0.weight"].view(E * r * r, E), p["ps_head.conv_expand.0.bias"])
    y = F.leaky_relu(y, 0.01)
    y = y.view(B, H, W, E, r, r).permute(0, 1, 4, 2, 5, 3).reshape(B, H * r, W * r, E)
    y = F.linear(y, p["decoder_pred.weight"].view(cfg.in_chans, E))
    return y.permute(0, 3, 1, 2)


def loss_fn(pred, target, cfg):
    """mean |pred-target| and, with log_transform, mean |expm1(pred)-expm1(target)|. tulip.py:690-700."""
    loss = (pred - target).abs().mean()
    if cfg.log_transform:
        pixel_loss = (torch.expm1(pred) - torch.expm1(target)).abs().mean()
    else:
        pixel_loss = loss.clone()
    return loss, pixel_loss


def drop_path_rates(cfg: Cfg):
    """Per-block stochastic-depth rates: encoder stage s uses linspace(0, rate, sum(depths))
    sliced per stage (tulip.py:409-410); decoder stage u reuses stage L-u-2 (tulip.py:447-453)."""
    dpr = [r.item() for r in torch.linspace(0, cfg.drop_path_rate, sum(cfg.depths))]
    enc = [dpr[sum(cfg.depths[:s]):sum(cfg.depths[:s + 1])] for s in range(cfg.num_layers)]
    dec = [enc[cfg.num_layers - u - 2] for u in range(cfg.num_layers - 1)]
    return enc, dec


# ----------------------------------------------------------------------------- whole model


def forward(p: dict, cfg: Cfg, x, target=None, state=None, drop_scales=None, taps=None):
    """TULIP.forward (tulip/model/tulip.py:702-737).

    drop_scales: optional dict {block prefix: (scale_attn, scale_mlp)} of per-sample DropPath
    scales for train-mode checks; None = eval / rate 0.
    taps: optional dict that receives named intermediate activations.
    Returns pred if target is None (mc_drop=True in the reference) else (pred, loss, pixel_loss).
    """
    Ls, win = cfg.num_layers, cfg.window_size
    ds = drop_scales or {}

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    x = patch_embed(x, p, cfg)
    tap("patch_embed", x)
    saves = []
    for s in range(Ls):
        saves.append(x)
        for b in range(cfg.depths[s]):
            pre = f"layers.{s}.blocks.{b}"
            x = swin_block(x, p, pre, cfg.num_heads[s], win, b % 2 == 1, cfg, state, ds.get(pre, (None, None)))
            tap(pre, x)
        if s < Ls - 1:
            x = patch_merging(x, p, f"layers.{s}.downsample", cfg)
            tap(f"layers.{s}.downsample", x)
    up = patch_unmerging if cfg.patch_unmerging else (lambda x_, p_, pre_: patch_expanding(x_, p_, pre_, cfg))
    x = up(x, p, "first_patch_expanding")
    tap("first_patch_expanding", x)
    for u in range(Ls - 1):
        s = Ls - u - 2
        x = torch.cat([x, saves[len(saves) - u - 2]], dim=-1)
        x = F.linear(x, p[f"skip_connection_layers.{u}.weight"], p[f"skip_connection_layers.{u}.bias"])
        tap(f"skip_connection_layers.{u}", x)
        for b in range(cfg.depths[s]):
            pre = f"layers_up.{u}.blocks.{b}"
            x = swin_block(x, p, pre, cfg.num_heads[s], win, b % 2 == 1, cfg, state, ds.get(pre, (None, None)))
            tap(pre, x)
        if u < Ls - 2:
            x = up(x, p, f"layers_up.{u}.upsample")
            tap(f"layers_up.{u}.upsample", x)
    pred = head(x, p, cfg) if cfg.pixel_shuffle else head_expanding(x, p, cfg)
    if target is None:
        return pred
    loss, pixel_loss = loss_fn(pred, target, cfg)
    return pred, loss, pixel_loss


def to_torch(params_np: dict, dtype=torch.float32, requires_grad=False, device="cpu") -> dict:
    out = {}
    for k, v in params_np.items():
        t = torch.from_numpy(v).to(device)
        if t.is_floating_point():
            t = t.to(dtype).requires_grad_(requires_grad)
        out[k] = t
    return out


def flops_per_frame(cfg: Cfg) -> dict:
    """2*MAC count of the matmul/conv work of one frame (forward), split by kernel family.
    Matches torch.utils.flop_counter on the reference (15 451 815 936 for tulip_base KITTI)."""
    E, Ls = cfg.embed_dim, cfg.num_layers
    H0, W0 = cfg.grid
    L = cfg.window_size[0] * cfg.window_size[1]
    tot = {"linear": 0, "attn": 0, "conv": 0}
    tot["conv"] += 2 * H0 * W0 * E * cfg.in_chans * cfg.patch_size[0] * 8

    def block(T, C):
        tot["linear"] += 2 * T * C * (3 * C + C + 4 * C + 4 * C)
        tot["attn"] += 2 * 2 * T * L * C

    for s in range(Ls):
        T, C = (H0 >> s) * (W0 >> s), E << s
        for _ in range(cfg.depths[s]):
            block(T, C)
        if s < Ls - 1:
            tot["linear"] += 2 * (T // 4) * 4 * C * 2 * C
    Tt, Ct = (H0 >> (Ls - 1)) * (W0 >> (Ls - 1)), E << (Ls - 1)
    tot["conv"] += 2 * Tt * Ct * 2 * Ct
    for u in range(Ls - 1):
        s = Ls - u - 2
        T, C = (H0 >> s) * (W0 >> s), E << s
        tot["linear"] += 2 * T * 2 * C * C
        for _ in range(cfg.depths[s]):
            block(T, C)
        if u < Ls - 2:
            tot["conv"] += 2 * T * C * 2 * C
    r2 = cfg.upscale_factor ** 2
    tot["conv"] += 2 * H0 * W0 * E * E * r2 + 2 * H0 * W0 * r2 * E * cfg.in_chans
    tot["total"] = sum(tot.values())
    return tot
