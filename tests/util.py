import os

import numpy as np
import torch

from oracle.params import bf16_bits_to_f32


def rel_l2(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf16r(t):
    """round an fp32 tensor to bf16 and back (what a bf16 store keeps)."""
    return t.to(torch.bfloat16).to(torch.float32)


def ulp_frac(a, b, ulps=1):
    """fraction of elements of `a` within `ulps` bf16 ulps of `b` (both fp32 holding bf16-representable values)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    ulp = torch.maximum(b.abs(), torch.tensor(1e-30)) * 2.0 ** -7       # bf16: 8 significand bits
    return float(((a - b).abs() <= ulps * ulp).float().mean())


def load_modules(golden_dir):
    return np.load(os.path.join(golden_dir, "modules.npz"))


def mod_params(mods, tag, device="cpu"):
    out, pre = {}, f"{tag}.p."
    for k in mods.files:
        if k.startswith(pre):
            v = mods[k]
            out[k[len(pre):]] = torch.from_numpy(bf16_bits_to_f32(v) if v.dtype == np.uint16 else v).to(device)
    return out


def dec(mods, key, device="cpu"):
    return torch.from_numpy(bf16_bits_to_f32(mods[key])).to(device)
