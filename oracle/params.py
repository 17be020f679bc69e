"""Parameter schema and deterministic parameter / input generators for the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).

``param_shapes`` restates the reference ``state_dict`` schema (SURVEY.md App. B;
reference constructor tulip/model/tulip.py:531-584, 643-688) so that tests can
build parameter dictionaries without instantiating either model.  The schema is
checked key-for-key against the real reference in ``oracle/make_golden.py``.

``make_params`` / ``make_inputs`` draw from numpy's PCG64 (bit-stable across numpy
versions and platforms), NOT from torch's RNG, so the fixtures under
``tests/golden`` can be regenerated bit-identically on any box.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np


@dataclass(frozen=True)
class Cfg:
    """Model configuration; defaults are the shipped `tulip_base` KITTI flags
    (bash_scripts/tulip_upsampling_kitti.sh:12-16,27-31; tulip.py:739-746)."""
    img_size: tuple = (16, 1024)
    target_img_size: tuple = (64, 1024)
    patch_size: tuple = (1, 4)
    in_chans: int = 1
    embed_dim: int = 96
    window_size: tuple = (2, 8)
    depths: tuple = (2, 2, 2, 2)
    num_heads: tuple = (3, 6, 12, 24)
    mlp_ratio: int = 4
    drop_path_rate: float = 0.1
    log_transform: bool = True
    ln_eps: float = 1e-6
    patch_unmerging: bool = True      # False: PatchExpanding (Linear + rearrange + LayerNorm, tulip.py:126-141) in the decoder
    pixel_shuffle: bool = True        # False: FinalPatchExpanding head (tulip.py:144-159) instead of PixelShuffleHead

    @property
    def num_layers(self) -> int:
        return len(self.depths)

    @property
    def grid(self) -> tuple:
        return (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])

    @property
    def upscale_factor(self) -> int:
        # tulip.py:577
        t, i, p = self.target_img_size, self.img_size, self.patch_size
        return int(((t[0] * t[1]) / (i[0] * i[1])) ** 0.5) * 2 * int(((p[0] * p[1]) // 4) ** 0.5)


TULIP_BASE = Cfg()
TULIP_LARGE = Cfg(depths=(2, 2, 2, 2, 2), num_heads=(3, 6, 12, 24, 48))


def _block_shapes(prefix: str, C: int, heads: int, win: tuple, mlp_ratio: int) -> list:
    nbias = (2 * win[0] - 1) * (2 * win[1] - 1)
    L = win[0] * win[1]
    return [
        (f"{prefix}.norm1.weight", (C,)), (f"{prefix}.norm1.bias", (C,)),
        (f"{prefix}.attn.relative_position_bias_table", (nbias, heads)),
        (f"{prefix}.attn.relative_position_index", (L, L)),          # int64 buffer
        (f"{prefix}.attn.qkv.weight", (3 * C, C)), (f"{prefix}.attn.qkv.bias", (3 * C,)),
        (f"{prefix}.attn.proj.weight", (C, C)), (f"{prefix}.attn.proj.bias", (C,)),
        (f"{prefix}.norm2.weight", (C,)), (f"{prefix}.norm2.bias", (C,)),
        (f"{prefix}.mlp.fc1.weight", (mlp_ratio * C, C)), (f"{prefix}.mlp.fc1.bias", (mlp_ratio * C,)),
        (f"{prefix}.mlp.fc2.weight", (C, mlp_ratio * C)), (f"{prefix}.mlp.fc2.bias", (C,)),
    ]


def param_shapes(cfg: Cfg) -> "OrderedDict[str, tuple]":
    """state_dict keys -> shapes in the reference's registration order
    (pos_drop, layers, layers_up, first_patch_expanding, skip_connection_layers,
    norm_up, patch_embed, decoder_pred, ps_head; tulip.py:553-580)."""
    E, Ls = cfg.embed_dim, cfg.num_layers
    out: list = []
    for s in range(Ls):                                   # encoder, tulip.py:643-660, 399-429
        C = E * 2 ** s
        for b in range(cfg.depths[s]):
            out += _block_shapes(f"layers.{s}.blocks.{b}", C, cfg.num_heads[s], cfg.window_size, cfg.mlp_ratio)
        if s < Ls - 1:                                    # PatchMerging, tulip.py:76-81
            out += [(f"layers.{s}.downsample.norm.weight", (4 * C,)),
                    (f"layers.{s}.downsample.norm.bias", (4 * C,)),
                    (f"layers.{s}.downsample.reduction.weight", (2 * C, 4 * C))]
    for u in range(Ls - 1):                               # decoder, tulip.py:662-680, 441-475
        s = Ls - u - 2
        C = E * 2 ** s
        for b in range(cfg.depths[s]):
            out += _block_shapes(f"layers_up.{u}.blocks.{b}", C, cfg.num_heads[s], cfg.window_size, cfg.mlp_ratio)
        if u < Ls - 2 and cfg.patch_unmerging:            # PatchUnmerging, tulip.py:109-115
            out += [(f"layers_up.{u}.upsample.expand.weight", (2 * C, C, 1, 1)),
                    (f"layers_up.{u}.upsample.expand.bias", (2 * C,))]
        elif u < Ls - 2:                                  # PatchExpanding, tulip.py:126-132
            out += [(f"layers_up.{u}.upsample.expand.weight", (2 * C, C)),
                    (f"layers_up.{u}.upsample.norm.weight", (C // 2,)),
                    (f"layers_up.{u}.upsample.norm.bias", (C // 2,))]
    Ctop = E * 2 ** (Ls - 1)
    if cfg.patch_unmerging:
        out += [("first_patch_expanding.expand.weight", (2 * Ctop, Ctop, 1, 1)),
                ("first_patch_expanding.expand.bias", (2 * Ctop,))]
    else:                                                 # tulip.py:565
        out += [("first_patch_expanding.expand.weight", (2 * Ctop, Ctop)),
                ("first_patch_expanding.norm.weight", (Ctop // 2,)),
                ("first_patch_expanding.norm.bias", (Ctop // 2,))]
    for u in range(Ls - 1):                               # skip Linear(2C->C), tulip.py:682-688
        C = E * 2 ** (Ls - 2 - u)
        out += [(f"skip_connection_layers.{u}.weight", (C, 2 * C)),
                (f"skip_connection_layers.{u}.bias", (C,))]
    out += [("norm_up.weight", (E,)), ("norm_up.bias", (E,))]
    out += [("patch_embed.proj.weight", (E, cfg.in_chans, cfg.patch_size[0], 8)),
            ("patch_embed.proj.bias", (E,)),
            ("patch_embed.norm.weight", (E,)), ("patch_embed.norm.bias", (E,))]
    out += [("decoder_pred.weight", (cfg.in_chans, E, 1, 1))]
    r2 = cfg.upscale_factor ** 2
    if cfg.pixel_shuffle:
        out += [("ps_head.conv_expand.0.weight", (E * r2, E, 1, 1)),
                ("ps_head.conv_expand.0.bias", (E * r2,))]
    else:                                                 # FinalPatchExpanding, tulip.py:144-150, 582
        out += [("final_patch_expanding.expand.weight", (E * r2, E)),
                ("final_patch_expanding.norm.weight", (E,)),
                ("final_patch_expanding.norm.bias", (E,))]
    return OrderedDict(out)


def make_params(cfg: Cfg, seed: int = 0, bf16_round: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Deterministic, well-conditioned parameters (NOT the reference init: biases and
    LayerNorm affine terms are made non-trivial so that every term is exercised).

    weights ~ N(0, 1/sqrt(fan_in))*0.7, biases ~ N(0, 0.05), LN weight ~ 1 + N(0, 0.1),
    bias table ~ N(0, 0.2).  With bf16_round the values are rounded to bf16
    (round-to-nearest-even) and stored as fp32, so that the CUDA path's bf16
    working copies are exact (parity protocol).
    """
    from .index_ops import relative_position_index
    rng = np.random.Generator(np.random.PCG64(seed))
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        if name.endswith("relative_position_index"):
            out[name] = relative_position_index(cfg.window_size)
            continue
        n = rng.standard_normal(shape).astype(np.float32)
        if name.endswith("relative_position_bias_table"):
            v = 0.2 * n
        elif ".norm" in name or name.startswith("norm_up"):
            v = 1.0 + 0.1 * n if name.endswith("weight") else 0.05 * n
        elif name.endswith("bias"):
            v = 0.05 * n
        else:
            fan_in = int(np.prod(shape[1:]))
            v = n * np.float32(0.7 / np.sqrt(fan_in))
        v = v.astype(np.float32)
        out[name] = round_bf16(v) if bf16_round else v
    return out


def make_inputs(cfg: Cfg, batch: int, seed: int = 1, bf16_round: bool = True, invalid_frac: float = 0.15):
    """Synthetic range images shaped like the real pipeline (datasets.py:68-70,143-150,285-294):
    log1p(U[0,1)) with `invalid_frac` of the pixels zeroed (invalid returns)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w = cfg.img_size
    H, W = cfg.target_img_size
    hi = np.log1p(rng.random((batch, cfg.in_chans, H, W), dtype=np.float32))
    hi = np.where(rng.random(hi.shape) < invalid_frac, np.float32(0), hi).astype(np.float32)
    step = H // h
    lo = np.ascontiguousarray(hi[:, :, ::step, :][:, :, :h, :w])        # row-downsampled view, like DownsampleTensor
    if bf16_round:
        lo, hi = round_bf16(lo), round_bf16(hi)
    return lo, hi


def round_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16 -> fp32 (pure integer arithmetic, bit-exact vs torch)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """bf16-exact fp32 array -> uint16 bit patterns (fixture storage at half the bytes)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    assert not np.any(u & 0xFFFF), "array is not bf16-exact"
    return (u >> 16).astype(np.uint16).reshape(x.shape)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32).reshape(b.shape)
