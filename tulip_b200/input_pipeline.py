"""Input pipeline on the GPU: the per-frame transform chains of the reference's datasets
(tulip/util/datasets.py:244-369) as one kernel that reads the raw range frames (metres, as the loaders return them) and writes
both model inputs.  The DataLoader then only has to deliver raw frames (e.g. pinned (B,H,W,2) float32 batches)."""
from __future__ import annotations

import torch

from ._lib import check, current_stream, load_library, ptr

# (scale, FilterInvalidPixels minimum or None): datasets.py:249-250 (durlar), :285-286 (kitti), :322-323 (carla)
DATASETS = {"kitti": (1 / 80, None), "durlar": (1 / 120, 0.3 / 120), "carla": (1 / 80, 2 / 80)}


def preprocess(raw: torch.Tensor, dataset: str, img_size_low_res, log_transform: bool = True):
    """raw [B,H,W,C] (channel 0 = range) or [B,H,W], fp32 CUDA, metres -> (lo [B,1,h,w], hi [B,1,H,W]) exactly as the reference's
    low-res / high-res datasets would deliver them (rows 0, f, 2f, ... of the same frame for the low-resolution input)."""
    if dataset not in DATASETS:
        raise NotImplementedError(f"Cannot find the dataset: {dataset}")
    if not raw.is_cuda:
        raise RuntimeError("tulip_b200.input_pipeline runs on CUDA only")
    x = raw.detach().to(torch.float32).contiguous()
    channels = x.shape[3] if x.dim() == 4 else 1
    B, H, W = x.shape[:3]
    h, w = img_size_low_res
    if H % h or W % w:
        raise ValueError(f"frame {H}x{W} is not a multiple of the low-resolution size {h}x{w}")
    scale, fmin = DATASETS[dataset]
    hi = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    lo = torch.empty((B, 1, h, w), dtype=torch.float32, device=x.device)
    check(load_library().tulip_preprocess_range(ptr(x), channels, float(scale), int(fmin is not None), float(fmin or 0.0), 1.0, H // h, W // w,
                                                int(log_transform), ptr(hi), ptr(lo), B, H, W, current_stream()), "tulip_preprocess_range")
    return lo, hi
