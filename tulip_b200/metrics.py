"""Evaluation metrics on the GPU: the step right after the model in the reference's `evaluate()`
(engine_upsampling.py:223-277), which runs on the host there (numpy projections, dense boolean voxel grids, a third-party Chamfer
extension).  Host side here: the sensor angle tables; device side: tulip_range_to_points / tulip_voxel_metrics /
tulip_chamfer_distance (csrc/metrics.cu)."""
from __future__ import annotations

import functools

import numpy as np
import torch

from ._lib import check, current_stream, load_library, ptr

MAX_RANGE = {"kitti": 80.0, "carla": 80.0}            # engine_upsampling.py:223-224, :232-233 (durlar: LUT projection, not built yet)


@functools.lru_cache(maxsize=16)
def _angle_tables_np(dataset: str, rows: int, cols: int):
    """float32 sines / cosines of the column (azimuth) and row (elevation) angles, computed with the arithmetic of
    evaluation.py:52-72 (kitti) / :90-104 (carla) so that the device products are bit-identical with the reference's."""
    if dataset == "kitti":
        ang_start_y, ang_res_y, ang_res_x = 24.8, 26.8 / (rows - 1), 360 / cols
        vertical = np.float32(np.arange(rows, dtype=np.float64) * ang_res_y) - ang_start_y
        horizon = -np.float32(np.arange(cols, dtype=np.float64) + 1 - (cols / 2)) * ang_res_x + 90.0
        vertical, horizon = vertical / 180.0 * np.pi, horizon / 180.0 * np.pi
    elif dataset == "carla":
        vertical = np.deg2rad(np.linspace(start=-15, stop=15, num=rows).astype(np.float32))
        horizon = np.deg2rad(np.linspace(start=-180, stop=180, num=cols, endpoint=False).astype(np.float32))
    else:
        raise NotImplementedError(f"range-image projection for dataset {dataset!r} is not built (kitti, carla)")
    return tuple(np.ascontiguousarray(t, dtype=np.float32) for t in (np.sin(horizon), np.cos(horizon), np.sin(vertical), np.cos(vertical)))


def angle_tables(dataset: str, rows: int, cols: int, device):
    cache = angle_tables.__dict__.setdefault("_dev", {})
    key = (dataset, rows, cols, str(device))
    if key not in cache:
        cache[key] = tuple(torch.from_numpy(t).to(device) for t in _angle_tables_np(dataset, rows, cols))
    return cache[key]


def range_to_points(img: torch.Tensor, dataset: str = "kitti", maximum_range: float | None = None) -> torch.Tensor:
    """img [B,1,H,W] or [B,H,W] or [H,W] normalised range (fp32, CUDA) -> points [B, H*W, 3] (img_to_pcd_kitti / img_to_pcd_carla)."""
    if not img.is_cuda:
        raise RuntimeError("tulip_b200.metrics runs on CUDA only")
    x = img.detach().to(torch.float32)
    if x.dim() == 4:
        x = x[:, 0]
    elif x.dim() == 2:
        x = x[None]
    x = x.contiguous()
    B, H, W = x.shape
    sh, ch, sv, cv = angle_tables(dataset, H, W, x.device)
    pts = torch.empty((B, H * W, 3), dtype=torch.float32, device=x.device)
    check(load_library().tulip_range_to_points(ptr(x), ptr(sh), ptr(ch), ptr(sv), ptr(cv), float(maximum_range or MAX_RANGE[dataset]),
                                               ptr(pts), B, H, W, current_stream()), "tulip_range_to_points")
    return pts


def voxel_metrics(pts_pred: torch.Tensor, pts_gt: torch.Tensor, grid_size: float = 0.1) -> torch.Tensor:
    """two clouds [n,3] fp32 -> float64 tensor {iou, precision, recall, f1} (device)."""
    a, b = pts_pred.detach().to(torch.float32).contiguous(), pts_gt.detach().to(torch.float32).contiguous()
    if a.shape != b.shape or a.dim() != 2 or a.shape[1] != 3:
        raise ValueError("voxel_metrics: clouds must both be [n, 3]")
    lib = load_library()
    ws = torch.empty(int(lib.tulip_voxel_metrics_workspace_bytes(a.shape[0])), dtype=torch.uint8, device=a.device)
    out = torch.empty(4, dtype=torch.float64, device=a.device)
    check(lib.tulip_voxel_metrics(ptr(a), ptr(b), a.shape[0], float(grid_size), ptr(ws), ptr(out), current_stream()), "tulip_voxel_metrics")
    return out


def chamfer_distance(points1: torch.Tensor, points2: torch.Tensor):
    """-> (cd = mean(dist1) + mean(dist2) as a device scalar, dist1 [n1], dist2 [n2]); squared nearest-neighbour distances."""
    a, b = points1.detach().to(torch.float32).contiguous(), points2.detach().to(torch.float32).contiguous()
    d1 = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    d2 = torch.empty(b.shape[0], dtype=torch.float32, device=a.device)
    out = torch.empty(3, dtype=torch.float32, device=a.device)
    check(load_library().tulip_chamfer_distance(ptr(a), ptr(b), a.shape[0], b.shape[0], ptr(d1), ptr(d2), ptr(out), current_stream()),
          "tulip_chamfer_distance")
    return out[0], d1, d2


def evaluate_frame(pred_img: torch.Tensor, gt_img: torch.Tensor, dataset: str = "kitti", grid_size: float = 0.1) -> dict:
    """The metric block of evaluate() for one frame (engine_upsampling.py:223-285): both range images (post-processed, linear
    normalised range, [H,W] or [1,1,H,W]) -> chamfer_dist, iou, precision, recall, f1 as Python floats (one D2H copy)."""
    pp, pg = range_to_points(pred_img, dataset)[0], range_to_points(gt_img, dataset)[0]
    cd, _, _ = chamfer_distance(pg, pp)                              # :259
    vm = voxel_metrics(pp, pg, grid_size)                            # :262-277
    vals = torch.cat([cd.to(torch.float64).reshape(1), vm]).cpu().tolist()
    return dict(zip(("chamfer_dist", "iou", "precision", "recall", "f1"), vals))
