"""Input pipeline on the GPU: the per-frame transform chains of the reference's datasets
(tulip/util/datasets.py:244-369) as one kernel that reads the raw range frames (metres, as the loaders return them) and writes
both model inputs.  The DataLoader then only has to deliver raw frames (e.g. pinned (B,H,W,2) float32 batches)."""
from __future__ import annotations

import numpy as np
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


def read_npy(paths, device="cuda"):
    """KITTI / DurLAR `.npy` frames (npy_loader, datasets.py:187-191: `np.load(f)[..., 0].astype(float32)`) -> one pinned upload of
    the raw (B, H, W, C) float32 batch; `preprocess` picks channel 0 on the GPU, so the host touches no pixel."""
    if isinstance(paths, str) or hasattr(paths, "__fspath__"):
        paths = [paths]
    frames = [np.load(p) for p in paths]
    shape = frames[0].shape
    if any(f.shape != shape for f in frames):
        raise ValueError("npy frames of one batch differ in shape")
    if frames[0].ndim not in (2, 3):
        raise ValueError(f"expected (H, W) or (H, W, C) range frames, got {shape}")
    host = torch.from_numpy(np.stack(frames).astype(np.float32, copy=False))
    return host.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else host


def read_rimg(paths, device="cuda"):
    """CARLA `.rimg` files (rimg_loader, datasets.py:181-193) -> frames [B, size0, size1] fp32 on the GPU, metres.  The host only
    strips the 16-byte headers and uploads the float16 payloads; transpose, flip and widening run in one kernel.  All files of a
    batch must have the same size (they do per resolution folder)."""
    if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__"):
        paths = [paths]
    size, rows = None, []
    for p in paths:
        buf = p if isinstance(p, (bytes, bytearray, memoryview)) else open(p, "rb").read()
        s = tuple(int(v) for v in np.frombuffer(buf, dtype=np.uint64, count=2))
        if size is None:
            size = s
        elif s != size:
            raise ValueError(f"rimg files of one batch differ in size: {s} vs {size}")
        body = np.frombuffer(buf, dtype=np.float16, offset=16)
        if body.size != s[0] * s[1]:
            raise ValueError(f"rimg payload holds {body.size} values, header says {s[0]} x {s[1]}")
        rows.append(body)
    host = torch.from_numpy(np.stack(rows))
    dev = host.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else host
    return decode_rimg(dev, size)


def decode_rimg(rows: torch.Tensor, size):
    """rows [B, size1 * size0] float16 CUDA (file payloads as stored) -> frames [B, size0, size1] fp32."""
    if not rows.is_cuda:
        raise RuntimeError("tulip_b200.input_pipeline runs on CUDA only")
    if rows.dtype != torch.float16:
        raise TypeError("rimg payloads are float16")
    s0, s1 = int(size[0]), int(size[1])
    rows = rows.contiguous().view(-1, s1 * s0)
    out = torch.empty((rows.shape[0], s0, s1), dtype=torch.float32, device=rows.device)
    check(load_library().tulip_rimg_decode(ptr(rows), ptr(out), rows.shape[0], s0, s1, current_stream()), "tulip_rimg_decode")
    return out
