"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the post-processing that `evaluate()` applies to a prediction before the
point-cloud metrics (reference tulip/engine_upsampling.py:174-244).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import this module; the product path (tulip_b200.ops.eval_postprocess -> CUDA) never does.

Pinned by tests/test_oracle_eval_post.py against a statement-by-statement torch execution of the reference lines (the reference
keeps them inline in evaluate(), so they cannot be imported)."""
import numpy as np

# engine_upsampling.py:183-188: lower bound of the valid normalised range per dataset (upper bound is 1 for all of them)
CLIP_LO = {"kitti": 2.0 / 80.0, "carla": 2.0 / 80.0, "durlar": 0.3 / 120.0}


def eval_postprocess(pred, lo, hi, log_transform=True, dataset="kitti"):
    """pred, hi: (B,1,H,W) float32; lo: (B,1,h,W).  Returns (out (B,1,H,W), losses (B,2)) with
    losses[:,0] = pixel loss (:192-193) and losses[:,1] = loss on the sensor's own rows (:216-219, :241-242; 0 when the rows are
    not kept: carla with different widths, :207-208)."""
    pred = np.asarray(pred, np.float32).copy()
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    if log_transform:                                            # :177-180
        pred, hi, lo = np.expm1(pred), np.expm1(hi), np.expm1(lo)
    cl = np.float32(CLIP_LO[dataset])
    pred = np.where((pred >= cl) & (pred <= np.float32(1.0)), pred, np.float32(0.0)).astype(np.float32)   # :183-188
    B, _, H, W = pred.shape
    h = lo.shape[2]
    losses = np.zeros((B, 2), np.float32)
    losses[:, 0] = np.abs(pred - hi).reshape(B, -1).mean(axis=1)                                       # :192-193
    keep = not (dataset == "carla" and lo.shape[3] != hi.shape[3])                                      # :207-208
    if keep:
        rows = np.arange(0, H, H // h)                                                                  # :214, :237
        part = pred[:, :, rows, :]
        losses[:, 1] = np.abs(part - lo).reshape(B, -1).mean(axis=1)                                    # :216-219
        pred[:, :, rows, :] = lo                                                                        # :221, :244
    return pred, losses


def mc_dropout_aggregate(preds, noise_threshold=0.03):
    """preds (N,1,H,W) float32 -> (1,1,H,W): engine_upsampling.py:423-427 (mean, unbiased std, zero where std > thr * mean)."""
    preds = np.asarray(preds, np.float32)
    mean = preds.mean(axis=0, keepdims=True, dtype=np.float32)
    std = preds.std(axis=0, keepdims=True, ddof=1, dtype=np.float32)
    out = mean.copy()
    out[std > np.float32(noise_threshold) * mean] = 0
    return out, std
