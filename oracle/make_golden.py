"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.npz.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container,
where /root/reference exists; the GPU box never executes this file (it consumes
the committed fixtures instead).

    python -m oracle.make_golden            # regenerate fixtures + print oracle-vs-reference deltas

The reference module tulip/model/tulip.py is imported as-is, with two import
stubs for packages that are not installed and not on the arithmetic path
(`chamfer_distance`, `timm.models.layers`).
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import index_ops, tulip_oracle as O
from .params import (Cfg, TULIP_BASE, TULIP_LARGE, f32_to_bf16_bits, make_inputs, make_params, param_shapes,
                     round_bf16)

REF_ROOT = os.environ.get("TULIP_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    """import model.tulip from the read-only reference tree with the two stubs."""
    if "chamfer_distance" not in sys.modules:
        cd = types.ModuleType("chamfer_distance")
        cd.ChamferDistance = object
        sys.modules["chamfer_distance"] = cd
    if "timm" not in sys.modules:
        timm, tm, tl = (types.ModuleType(n) for n in ("timm", "timm.models", "timm.models.layers"))

        class DropPath(nn.Module):                    # only referenced by the (dead) Swin-V2 file
            def __init__(self, p=0.0):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x

        tl.DropPath = DropPath
        tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        tl.trunc_normal_ = nn.init.trunc_normal_
        timm.models, tm.layers = tm, tl
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    sys.path.insert(0, os.path.join(REF_ROOT, "tulip"))
    import model.tulip as T
    return T


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def build_reference(T, cfg: Cfg, large: bool):
    fn = T.tulip_large if large else T.tulip_base
    return fn(img_size=tuple(cfg.img_size), target_img_size=tuple(cfg.target_img_size),
              patch_size=tuple(cfg.patch_size), in_chans=cfg.in_chans, window_size=list(cfg.window_size),
              swin_v2=False, pixel_shuffle=True, circular_padding=True, log_transform=cfg.log_transform,
              patch_unmerging=True)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def model_fixture(T, name: str, cfg: Cfg, large: bool, batch: int, pseed: int, xseed: int, store_pred_stride: int = 1):
    """Whole-model golden: reference fwd+bwd in fp32/eval on PCG64 params; the fixture keeps the
    outputs, a strided sample of pred, the losses and per-parameter gradient norms + samples."""
    ref = build_reference(T, cfg, large).eval()
    sd = ref.state_dict()
    shapes = param_shapes(cfg)
    assert list(sd.keys()) == list(shapes.keys()), "state_dict schema drift"
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    pn = make_params(cfg, pseed)
    ref.load_state_dict({k: torch.from_numpy(v) for k, v in pn.items()}, strict=True)
    lo, hi = make_inputs(cfg, batch, xseed)
    lo_t, hi_t = torch.from_numpy(lo), torch.from_numpy(hi)
    pred, loss, pixel = ref(lo_t, hi_t, eval=True)
    loss.backward()
    ref_grads = {k: p_.grad.detach() for k, p_ in ref.named_parameters()}

    # oracle on the same values
    po = O.to_torch(pn, requires_grad=True)
    pred_o, loss_o, pixel_o = O.forward(po, cfg, lo_t, hi_t, state={})
    loss_o.backward()
    worst = 0.0
    wk = None
    for k, g in ref_grads.items():
        r = rel(po[k].grad.detach(), g)
        if r > worst:
            worst, wk = r, k
    print(f"[{name}] oracle vs reference: pred rel {rel(pred_o.detach(), pred.detach()):.2e}  loss {abs(loss_o.item()-loss.item()):.2e} "
          f"pixel {abs(pixel_o.item()-pixel.item()):.2e}  worst grad rel {worst:.2e} ({wk})")
    # fp32 re-association noise only; the L1 sign() gradient makes LayerNorm-bias grads ill-conditioned
    assert rel(pred_o.detach(), pred.detach()) < 1e-5 and worst < 2e-3

    out = {
        "pred": pred.detach().numpy()[..., ::store_pred_stride].astype(np.float32),
        "pred_stride": np.int64(store_pred_stride),
        "pred_sum": np.float64(pred.double().sum().item()),
        "pred_sumsq": np.float64((pred.double() ** 2).sum().item()),
        "loss": np.float64(loss.item()),
        "pixel_loss": np.float64(pixel.item()),
        "batch": np.int64(batch), "pseed": np.int64(pseed), "xseed": np.int64(xseed),
        "param_sha": np.array(sha16(np.concatenate([v.reshape(-1).astype(np.float64) for v in pn.values()]))),
        "input_sha": np.array(sha16(lo) + sha16(hi)),
        "grad_names": np.array(list(ref_grads.keys())),
        "grad_norm": np.array([g.double().norm().item() for g in ref_grads.values()], dtype=np.float64),
        # first 64 entries of every grad, for element-level checks without shipping 108 MB
        "grad_head": np.stack([np.pad(g.reshape(-1)[:64].numpy(), (0, max(0, 64 - g.numel()))) for g in ref_grads.values()]),
    }
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)


def module_fixtures(T):
    """Per-module goldens from the reference's own classes on small seeded inputs."""
    g = np.random.Generator(np.random.PCG64(7))
    out = {}

    def rnd(*shape, s=1.0):
        return torch.from_numpy(round_bf16((g.standard_normal(shape) * s).astype(np.float32)))

    def enc(t):                                        # inputs / params are bf16-exact: store the 16-bit patterns
        return f32_to_bf16_bits(t.detach().numpy() if isinstance(t, torch.Tensor) else t)

    def fill(mod):
        sd = {}
        for k, v in mod.state_dict().items():
            if v.is_floating_point():
                t = rnd(*v.shape, s=0.2)
                if k.endswith("norm.weight") or k.endswith("norm1.weight") or k.endswith("norm2.weight"):
                    t = torch.from_numpy(round_bf16((t + 1).numpy()))
                sd[k] = t
            else:
                sd[k] = v
        mod.load_state_dict(sd)
        return {k: (enc(v) if v.is_floating_point() else v.numpy()) for k, v in sd.items()}

    norm = lambda c: nn.LayerNorm(c, eps=1e-6)
    # WindowAttention, unshifted and shifted, C=96 on an (2, 8, 32) grid
    for shift in (False, True):
        m = T.WindowAttention(96, window_size=[2, 8], num_heads=3, shift=shift).eval()
        sd = fill(m)
        x = rnd(1, 4, 16, 96).requires_grad_(True)
        y = m(x)
        gy = rnd(*y.shape)
        y.backward(gy)
        tag = f"attn_shift{int(shift)}"
        out.update({f"{tag}.x": enc(x), f"{tag}.y": y.detach().numpy(), f"{tag}.gy": enc(gy),
                    f"{tag}.gx": x.grad.numpy(),
                    f"{tag}.g_table": m.relative_position_bias_table.grad.numpy(),
                    f"{tag}.g_qkv_w": m.qkv.weight.grad.numpy()})
        out.update({f"{tag}.p.{k}": v for k, v in sd.items()})
    # backup-window path: H=1 < win_h
    m = T.WindowAttention(96, window_size=[2, 8], num_heads=3, shift=True).eval()
    sd = fill(m)
    x = rnd(2, 1, 32, 96)
    y = m(x)
    assert m.window_size == (1, 16) and m.shift_size == (0, 8)
    out.update({"attn_backup.x": enc(x), "attn_backup.y": y.detach().numpy()})
    out.update({f"attn_backup.p.{k}": v for k, v in sd.items()})
    # SwinTransformerBlock (shifted)
    m = T.SwinTransformerBlock(96, 3, window_size=[2, 8], shift=True, mlp_ratio=4, norm_layer=norm).eval()
    sd = fill(m)
    x = rnd(1, 4, 16, 96)
    out.update({"block.x": enc(x), "block.y": m(x).detach().numpy()})
    out.update({f"block.p.{k}": v for k, v in sd.items()})
    # PatchEmbedding
    m = T.PatchEmbedding(img_size=(4, 64), patch_size=(1, 4), in_c=1, embed_dim=96, norm_layer=norm, circular_padding=True).eval()
    sd = fill(m)
    x = rnd(2, 1, 4, 64)
    out.update({"embed.x": enc(x), "embed.y": m(x).detach().numpy()})
    out.update({f"embed.p.{k}": v for k, v in sd.items()})
    # PatchMerging / PatchUnmerging
    m = T.PatchMerging(96, norm_layer=norm).eval()
    sd = fill(m)
    x = rnd(2, 4, 8, 96)
    out.update({"merge.x": enc(x), "merge.y": m(x).detach().numpy()})
    out.update({f"merge.p.{k}": v for k, v in sd.items()})
    m = T.PatchUnmerging(192).eval()
    sd = fill(m)
    x = rnd(2, 2, 4, 192)
    out.update({"unmerge.x": enc(x), "unmerge.y": m(x).detach().numpy()})
    out.update({f"unmerge.p.{k}": v for k, v in sd.items()})
    # PixelShuffleHead + decoder_pred on a tiny grid
    ph = T.PixelShuffleHead(96, 4).eval()
    sdh = fill(ph)
    dp = nn.Conv2d(96, 1, kernel_size=(1, 1), bias=False)
    wd = rnd(1, 96, 1, 1, s=0.2)
    dp.weight.data.copy_(wd)
    x = rnd(2, 96, 2, 8)
    out.update({"head.x_nchw": enc(x), "head.y": dp(ph(x)).detach().numpy(), "head.wd": enc(wd)})
    out.update({f"head.p.{k}": v for k, v in sdh.items()})
    np.savez_compressed(os.path.join(GOLDEN, "modules.npz"), **out)
    print(f"[modules] wrote {len(out)} arrays")


def index_fixtures(T):
    """Known answers for the integer / index ops, straight from the reference."""
    out = {}
    wa = T.WindowAttention(dim=96, window_size=(2, 8), num_heads=3, shift=True)
    out["rel_index_2x8"] = wa.relative_position_index.numpy()
    x = torch.arange(16 * 256, dtype=torch.float32).view(1, 16, 256, 1)
    out["partition_16x256"] = wa.window_partition(x).numpy()
    out["mask_16x256"] = wa.create_mask(x).numpy()
    x2 = torch.zeros(1, 32, 512, 1)
    out["mask_32x512"] = wa.create_mask(x2).numpy()
    wb = T.WindowAttention(dim=96, window_size=(2, 8), num_heads=3, shift=True)
    wb.window_size, wb.shift_size = wb.backup_window_size, wb.backup_shift_size
    out["mask_backup_1x16"] = wb.create_mask(torch.zeros(1, 1, 16, 1)).numpy()
    out["mask_backup_1x64"] = wb.create_mask(torch.zeros(1, 1, 64, 1)).numpy()
    out["merge_4x4"] = T.PatchMerging.merging(torch.arange(16.).view(1, 4, 4, 1)).numpy()
    out["pixel_shuffle_r2"] = nn.PixelShuffle(2)(torch.arange(8.).view(1, 8, 1, 1)).numpy()
    out["pixel_shuffle_r4"] = nn.PixelShuffle(4)(torch.arange(2 * 32 * 2 * 3, dtype=torch.float32).view(2, 32, 2, 3)).numpy()
    pe = T.PatchEmbedding(img_size=(1, 1024), patch_size=(1, 4), in_c=1, embed_dim=96, circular_padding=True)
    out["circ_pad_1024"] = pe.circularpadding(torch.arange(1024.).view(1, 1, 1, 1024)).numpy()
    out["roll_m1_m4"] = torch.roll(torch.arange(4 * 16.).view(1, 4, 16, 1), shifts=(-1, -4), dims=(1, 2)).numpy()
    base = build_reference(T, TULIP_BASE, False)
    out["upscale_factor_kitti"] = np.int64(base.upscale_factor)
    out["grid_kitti"] = np.array(base.patch_embed.grid_size)
    enc = [[blk.drop_path.drop_prob if hasattr(blk.drop_path, "drop_prob") else 0.0 for blk in l.blocks] for l in base.layers]
    dec = [[blk.drop_path.drop_prob if hasattr(blk.drop_path, "drop_prob") else 0.0 for blk in l.blocks] for l in base.layers_up]
    out["drop_rates_enc"] = np.array(enc, dtype=np.float64)
    out["drop_rates_dec"] = np.array(dec, dtype=np.float64)
    for k in ("rel_index_2x8", "partition_16x256", "mask_16x256"):
        print(f"[index] {k}: sha {sha16(out[k])}")
    assert sha16(out["rel_index_2x8"]) == "4ddf27b1d7ef65c1"
    assert sha16(out["partition_16x256"]) == "18cb13e7e85d8ed1"
    assert sha16(out["mask_16x256"]) == "e85f862a06f42493"
    np.savez_compressed(os.path.join(GOLDEN, "index_ops.npz"), **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN, exist_ok=True)
    T = import_reference()
    index_fixtures(T)
    module_fixtures(T)
    model_fixture(T, "model_base_kitti_b2", TULIP_BASE, False, batch=2, pseed=0, xseed=1)
    durlar = Cfg(img_size=(32, 2048), target_img_size=(128, 2048))
    model_fixture(T, "model_base_durlar_b1", durlar, False, batch=1, pseed=2, xseed=3, store_pred_stride=8)
    model_fixture(T, "model_large_kitti_b1", TULIP_LARGE, True, batch=1, pseed=4, xseed=5, store_pred_stride=4)


if __name__ == "__main__":
    main()
