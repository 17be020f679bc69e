"""TEST INFRASTRUCTURE ONLY -- torch (CPU, fp32) restatement of the reference's per-frame input transforms
(tulip/util/datasets.py: npy_loader :187-191, rimg_loader :181-193 [the CARLA .rimg container], ToTensor, ScaleTensor :140-144, FilterInvalidPixels :146-154, DownsampleTensor
:120-128, DownsampleTensorWidth :130-138, LogTransform :73-75) chained as build_{kitti,durlar,carla}_upsampling_dataset do
(:244-369).  Only tests/, smoke() and bench.py's cpu_baseline may import this module.

Pinned: `python -m oracle.make_golden_input` (build container only) composes the UNMODIFIED reference transform classes the way the
three builders do and checks this file bit for bit, writing tests/golden/input_pipeline.npz."""
import numpy as np
import torch

# (scale, filter min or None) per dataset: datasets.py:249-250 (durlar), :285-286 (kitti), :322-323 (carla)
DATASETS = {"kitti": (1 / 80, None), "durlar": (1 / 120, 0.3 / 120), "carla": (1 / 80, 2 / 80)}


def preprocess(raw, dataset, h_low, w_low=None, log_transform=True):
    """raw: numpy / tensor (B, H, W, C) or (B, H, W) fp32 metres -> (lo (B,1,h_low,w_low), hi (B,1,H,W)) fp32."""
    x = torch.as_tensor(raw, dtype=torch.float32)
    if x.dim() == 4:
        x = x[..., 0]                                            # npy_loader keeps channel 0 (:189-190)
    x = x[:, None]                                               # ToTensor on a float (H, W) array: (1, H, W), no rescaling
    scale, fmin = DATASETS[dataset]
    x = x * scale                                                # ScaleTensor
    if fmin is not None:
        x = torch.where((x >= fmin) & (x <= 1), x, 0)            # FilterInvalidPixels(min_range, max_range = 1)
    H, W = x.shape[-2:]
    w_low = W if w_low is None else w_low
    lo = x[:, :, range(0, H, H // h_low), :]                     # DownsampleTensor(index 0)
    if W // w_low > 1:
        lo = lo[:, :, :, range(0, W, W // w_low)]                # DownsampleTensorWidth
    hi = x
    if log_transform:
        lo, hi = torch.log1p(lo), torch.log1p(hi)                # LogTransform
    return lo.contiguous(), hi.contiguous()


def rimg_decode(buf: bytes):
    """The CARLA `.rimg` container (datasets.py:181-193): two native unsigned longs (size[0], size[1]), then size[0] * size[1]
    float16 values stored as `size[1]` rows of `size[0]`; the loader transposes, flips BOTH axes and widens to float32.
    -> (frame (size[0], size[1]) float32, payload (size[1], size[0]) float16 as stored)."""
    size = np.frombuffer(buf, dtype=np.uint64, count=2)          # np.uint is the platform's unsigned long: 8 bytes on Linux
    s0, s1 = int(size[0]), int(size[1])
    payload = np.frombuffer(buf, dtype=np.float16, offset=16).reshape(s1, s0)
    frame = np.empty((s0, s1), dtype=np.float32)
    for i in range(s0):                                          # F[i, j] = payload[s1 - 1 - j, s0 - 1 - i]
        frame[i, :] = payload[::-1, s0 - 1 - i].astype(np.float32)
    return frame, payload
