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


@torch.no_grad()
def mc_dropout_upsample(model, x_lo: torch.Tensor, target: torch.Tensor, iterations: int = 50, iteration_batch: int = 8,
                        noise_threshold: float = 0.03):
    """MCdrop()'s prediction for ONE frame (engine_upsampling.py:365-427): `iterations` stochastic passes in batches of
    `iteration_batch` tiles of the input through `model(tile, target, mc_drop=True)`, then mean / std / noise removal as one
    kernel.  (Every Dropout of both factories has p = 0 and DropPath is not re-enabled by enable_dropout, so with the shipped
    models the passes are identical and std = 0 -- SURVEY 3.4; the path is the reference's nevertheless.)"""
    if x_lo.shape[0] != 1:
        raise ValueError("mc_dropout_upsample: one frame at a time (the reference's evaluation loader has batch_size 1)")
    if iterations <= iteration_batch:
        raise ValueError("iterations must exceed iteration_batch")                 # engine_upsampling.py:369
    preds = torch.empty((iterations, *target.shape[1:]), dtype=torch.float32, device=x_lo.device)
    done = 0
    while done < iterations:
        nb = min(iteration_batch, iterations - done)                               # :413
        preds[done:done + nb] = model(x_lo.expand(nb, -1, -1, -1).contiguous(), target, mc_drop=True)      # :414-421
        done += nb
    return ops.mc_dropout_aggregate(preds, noise_threshold)                        # :423-427
