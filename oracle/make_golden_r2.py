"""Round-2 additions to the golden fixtures, generated from the UNMODIFIED reference (build container only):

    python -m oracle.make_golden_r2

  tests/golden/modules_r2.npz         skip Linear on cat([x, skip]) (tulip.py:715-716), PatchUnmerging and the head
                                      (norm_up + PixelShuffleHead + decoder_pred + L1, tulip.py:720-731, 690-693) WITH their
                                      gradients, and the head at embed_dim 192
  tests/golden/model_wide_kitti_b1.npz   BASELINE cfg5 surrogate (SURVEY 8d option i): TULIP(depths=(2,2,18,2), embed_dim=192,
                                      num_heads=(6,12,24,48)) at 16x1024 -> 64x1024, batch 1, forward + backward

TEST INFRASTRUCTURE (see oracle/__init__.py).  The round-1 fixtures are not touched (own file, own PCG64 stream)."""
from __future__ import annotations

import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import tulip_oracle as O
from .make_golden import GOLDEN, import_reference, model_fixture, rel
from .params import Cfg, f32_to_bf16_bits, round_bf16

TULIP_WIDE = Cfg(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48))


def module_fixtures_r2(T):
    g = np.random.Generator(np.random.PCG64(11))
    out = {}

    def rnd(*shape, s=1.0):
        return torch.from_numpy(round_bf16((g.standard_normal(shape) * s).astype(np.float32)))

    def enc(t):
        return f32_to_bf16_bits(t.detach().numpy() if isinstance(t, torch.Tensor) else t)

    def fill(mod, s=0.2):
        sd = {}
        for k, v in mod.state_dict().items():
            t = rnd(*v.shape, s=s)
            if k.endswith("norm.weight") or k == "weight" and v.ndim == 1:
                t = torch.from_numpy(round_bf16((t + 1).numpy()))
            sd[k] = t
        mod.load_state_dict(sd)
        return {k: enc(v) for k, v in sd.items()}

    # skip connection: Linear(2C -> C) on cat([x, skip], -1)  (tulip.py:715-716, 682-688), C = 192 on a (4, 16) grid
    lin = nn.Linear(384, 192)
    sd = fill(lin, s=0.08)
    x, skip = rnd(2, 4, 16, 192).requires_grad_(True), rnd(2, 4, 16, 192).requires_grad_(True)
    y = lin(torch.cat([x, skip], -1))
    gy = rnd(*y.shape)
    y.backward(gy)
    out.update({"skip.x": enc(x), "skip.skip": enc(skip), "skip.y": y.detach().numpy(), "skip.gy": enc(gy), "skip.gx": x.grad.numpy(),
                "skip.gskip": skip.grad.numpy(), "skip.g_w": lin.weight.grad.numpy(), "skip.g_b": lin.bias.grad.numpy()})
    out.update({f"skip.p.{k}": v for k, v in sd.items()})

    # PatchUnmerging with gradients (tulip.py:109-123), C = 192 -> 96 on a (4, 8) -> (8, 16) grid
    m = T.PatchUnmerging(192)
    sd = fill(m, s=0.1)
    x = rnd(2, 4, 8, 192).requires_grad_(True)
    y = m(x)
    gy = rnd(*y.shape)
    y.backward(gy)
    out.update({"unmerge.x": enc(x), "unmerge.y": y.detach().numpy(), "unmerge.gy": enc(gy), "unmerge.gx": x.grad.numpy(),
                "unmerge.g_w": m.expand.weight.grad.numpy(), "unmerge.g_b": m.expand.bias.grad.numpy()})
    out.update({f"unmerge.p.{k}": v for k, v in sd.items()})

    # head: norm_up -> NCHW -> PixelShuffleHead -> decoder_pred -> mean |pred - target|   (tulip.py:720-731, 692-693)
    for E in (96, 192):
        tag = f"head{E}"
        norm = nn.LayerNorm(E, eps=1e-6)
        ph = T.PixelShuffleHead(E, 4)
        dp = nn.Conv2d(E, 1, kernel_size=(1, 1), bias=False)
        sdn, sdh, sdd = fill(norm, s=0.1), fill(ph, s=0.1), fill(dp, s=0.2)
        x = rnd(2, 4, 32, E).requires_grad_(True)                           # NHWC tokens, as the decoder hands them over
        pred = dp(ph(norm(x).permute(0, 3, 1, 2).contiguous()))
        target = rnd(*pred.shape, s=0.5)
        loss = (pred - target).abs().mean()
        loss.backward()
        out.update({f"{tag}.x": enc(x), f"{tag}.pred": pred.detach().numpy(), f"{tag}.target": enc(target),
                    f"{tag}.loss": np.float64(loss.item()), f"{tag}.gx": x.grad.numpy(),
                    f"{tag}.g_norm_w": norm.weight.grad.numpy(), f"{tag}.g_norm_b": norm.bias.grad.numpy(),
                    # (1536 | 3072, E, 1, 1): every row at E = 96, every 8th row at E = 192 (keeps the fixture small)
                    f"{tag}.g_we": ph.conv_expand[0].weight.grad.numpy()[::(1 if E == 96 else 8)], f"{tag}.g_be": ph.conv_expand[0].bias.grad.numpy(),
                    f"{tag}.g_wd": dp.weight.grad.numpy()})
        out.update({f"{tag}.norm.{k}": v for k, v in sdn.items()})
        out.update({f"{tag}.ps.{k}": v for k, v in sdh.items()})
        out.update({f"{tag}.dec.{k}": v for k, v in sdd.items()})
    np.savez_compressed(os.path.join(GOLDEN, "modules_r2.npz"), **out)
    print(f"[modules_r2] wrote {len(out)} arrays")


def wide_model_fixture(T):
    """model_fixture() builds through the factories; the surrogate needs the TULIP constructor itself (tulip.py:531-584)."""
    import oracle.make_golden as MG
    cfg = TULIP_WIDE

    def build_wide(T_, cfg_, large):
        return T_.TULIP(img_size=tuple(cfg_.img_size), target_img_size=tuple(cfg_.target_img_size), patch_size=tuple(cfg_.patch_size),
                        in_chans=cfg_.in_chans, embed_dim=cfg_.embed_dim, window_size=list(cfg_.window_size), depths=cfg_.depths,
                        num_heads=cfg_.num_heads, mlp_ratio=4, qkv_bias=True, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.1,
                        norm_layer=partial(nn.LayerNorm, eps=1e-6), swin_v2=False, pixel_shuffle=True, circular_padding=True,
                        log_transform=cfg_.log_transform, patch_unmerging=True)
    saved = MG.build_reference
    MG.build_reference = build_wide
    try:
        model_fixture(T, "model_wide_kitti_b1", cfg, False, batch=1, pseed=6, xseed=7, store_pred_stride=4)
    finally:
        MG.build_reference = saved


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    T = import_reference()
    module_fixtures_r2(T)
    wide_model_fixture(T)


if __name__ == "__main__":
    main()
