"""Evaluation / inference path: forward-only call of the drop-in module plus the post-processing that the reference's
`evaluate()` applies before its point-cloud metrics (engine_upsampling.py:168-244), as one fused CUDA kernel.  The forward is replayed
as a CUDA graph from the third identical call on (tulip_net::run_graphed), so the B = 1 loop of the reference's evaluation costs two
host calls per frame."""
from __future__ import annotations

import torch

from . import ops

# engine_upsampling.py:183-188: valid normalised range is [CLIP_LO, 1], everything else becomes 0 (no return)
CLIP_LO = {"kitti": 2.0 / 80.0, "carla": 2.0 / 80.0, "durlar": 0.3 / 120.0}


@torch.no_grad()
def upsample(model, x_lo: torch.Tensor, target: torch.Tensor, dataset: str = "kitti", log_transform: bool = True):
    """-> (range image [B,1,H,W] in linear normalised range with the sensor's own rows restored,
           losses [B,2] = per-frame {pixel loss (:192-193), loss on the sensor rows (:216-219)})."""
    if dataset not in CLIP_LO:
        raise NotImplementedError(f"Cannot find the dataset: {dataset}")          # engine_upsampling.py:253-254
    was_training = model.training
    if was_training:                                                               # (walking the module tree costs ~0.2 ms: call
        model.eval()                                                               #  model.eval() once yourself in a loop) :137
    try:
        with torch.no_grad():                                                      # evaluate() is @torch.no_grad() (:126): forward-only kernels
            pred, _, _ = model(x_lo, target, eval=True)                            # :169-171
    finally:
        if was_training:
            model.train(True)
    keep = not (dataset == "carla" and x_lo.shape[-1] != target.shape[-1])         # :207-208
    return ops.eval_postprocess(pred, x_lo, target, log_transform, CLIP_LO[dataset], keep)
