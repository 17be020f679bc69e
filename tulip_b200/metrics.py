"""Evaluation metrics on the GPU: the step right after the model in the reference's `evaluate()`
(engine_upsampling.py:223-277), which runs on the host there (numpy projections, dense boolean voxel grids, a third-party Chamfer
extension).  Host side here: the sensor angle tables; device side: tulip_range_to_points / tulip_voxel_metrics /
tulip_chamfer_distance (csrc/metrics.cu)."""
from __future__ import annotations

import functools

import numpy as np
import torch

from ._lib import check, current_stream, load_library, ptr

MAX_RANGE = {"kitti": 80.0, "carla": 80.0, "durlar": 120.0}          # engine_upsampling.py:223-224, :232-233, :250-251


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
        raise NotImplementedError(f"Cannot find the dataset: {dataset}")
    return tuple(np.ascontiguousarray(t, dtype=np.float32) for t in (np.sin(horizon), np.cos(horizon), np.sin(vertical), np.cos(vertical)))


# Ouster OS1-128 calibration of the DurLAR sensor (evaluation.py:7-17): per-row column stagger, beam elevations in degrees, beam-origin
# offsets.  px_to_xyz / idx_from_px (:19-45) are evaluated on the host for the per-column / per-row factors, in float64.
_DURLAR_OFFSET_LUT = np.tile(np.array([48, 32, 16, 0], dtype=np.int32), 32)
_DURLAR_ELEVATION_DEG = np.array([
    21.42, 21.12, 20.81, 20.5, 20.2, 19.9, 19.58, 19.26, 18.95, 18.65, 18.33, 18.02, 17.68, 17.37, 17.05, 16.73, 16.4, 16.08, 15.76, 15.43,
    15.1, 14.77, 14.45, 14.11, 13.78, 13.45, 13.13, 12.79, 12.44, 12.12, 11.77, 11.45, 11.1, 10.77, 10.43, 10.1, 9.74, 9.4, 9.06, 8.72,
    8.36, 8.02, 7.68, 7.34, 6.98, 6.63, 6.29, 5.95, 5.6, 5.25, 4.9, 4.55, 4.19, 3.85, 3.49, 3.15, 2.79, 2.44, 2.1, 1.75, 1.38, 1.03, 0.68,
    0.33, -0.03, -0.38, -0.73, -1.07, -1.45, -1.8, -2.14, -2.49, -2.85, -3.19, -3.54, -3.88, -4.26, -4.6, -4.95, -5.29, -5.66, -6.01,
    -6.34, -6.69, -7.05, -7.39, -7.73, -8.08, -8.44, -8.78, -9.12, -9.45, -9.82, -10.16, -10.5, -10.82, -11.19, -11.52, -11.85, -12.18,
    -12.54, -12.87, -13.2, -13.52, -13.88, -14.21, -14.53, -14.85, -15.2, -15.53, -15.84, -16.16, -16.5, -16.83, -17.14, -17.45, -17.8,
    -18.11, -18.42, -18.72, -19.06, -19.37, -19.68, -19.97, -20.31, -20.61, -20.92, -21.22])
_DURLAR_ORIGIN_OFFSET, _DURLAR_Z_OFFSET, _DURLAR_ANGLE_OFF = 0.015806, 0.03618, np.pi * 4.2285 / 180.


def durlar_tables(rows: int, cols: int, device):
    cache = durlar_tables.__dict__.setdefault("_dev", {})
    key = (rows, cols, str(device))
    if key not in cache:
        if rows > len(_DURLAR_ELEVATION_DEG):
            raise ValueError("the DurLAR sensor has 128 beams")
        u = (cols + np.arange(cols)) % cols
        encoder = 2.0 * np.pi - (u * (np.pi * 2.0 / cols))
        elevation = np.pi * _DURLAR_ELEVATION_DEG[:rows] / 180.
        f64 = [np.cos(encoder + _DURLAR_ANGLE_OFF), np.sin(encoder + _DURLAR_ANGLE_OFF), np.cos(encoder), np.sin(encoder), np.cos(elevation),
               np.sin(elevation)]
        cache[key] = tuple(torch.from_numpy(np.ascontiguousarray(t)).to(device) for t in f64) + (torch.from_numpy(_DURLAR_OFFSET_LUT[:rows].copy()).to(device),)
    return cache[key]


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
    if dataset == "durlar":                                         # Ouster LUT projection, float64 points (evaluation.py:19-58)
        ca, sa, ce, se, cel, sel, off = durlar_tables(H, W, x.device)
        pts64 = torch.empty((B, H * W, 3), dtype=torch.float64, device=x.device)
        check(load_library().tulip_range_to_points_durlar(ptr(x), ptr(ca), ptr(sa), ptr(ce), ptr(se), ptr(cel), ptr(sel), ptr(off),
                                                          float(maximum_range or MAX_RANGE[dataset]), _DURLAR_ORIGIN_OFFSET, _DURLAR_Z_OFFSET,
                                                          ptr(pts64), B, H, W, current_stream()), "tulip_range_to_points_durlar")
        return pts64
    sh, ch, sv, cv = angle_tables(dataset, H, W, x.device)
    pts = torch.empty((B, H * W, 3), dtype=torch.float32, device=x.device)
    check(load_library().tulip_range_to_points(ptr(x), ptr(sh), ptr(ch), ptr(sv), ptr(cv), float(maximum_range or MAX_RANGE[dataset]),
                                               ptr(pts), B, H, W, current_stream()), "tulip_range_to_points")
    return pts


def voxel_metrics(pts_pred: torch.Tensor, pts_gt: torch.Tensor, grid_size: float = 0.1) -> torch.Tensor:
    """two clouds [n,3] fp32 -> float64 tensor {iou, precision, recall, f1} (device)."""
    f64 = pts_pred.dtype == torch.float64                           # durlar clouds are float64 in the reference, kitti / carla float32
    dt = torch.float64 if f64 else torch.float32
    a, b = pts_pred.detach().to(dt).contiguous(), pts_gt.detach().to(dt).contiguous()
    if a.shape != b.shape or a.dim() != 2 or a.shape[1] != 3:
        raise ValueError("voxel_metrics: clouds must both be [n, 3]")
    lib = load_library()
    ws = torch.empty(int(lib.tulip_voxel_metrics_workspace_bytes(a.shape[0])), dtype=torch.uint8, device=a.device)
    out = torch.empty(4, dtype=torch.float64, device=a.device)
    fn = lib.tulip_voxel_metrics_f64 if f64 else lib.tulip_voxel_metrics
    check(fn(ptr(a), ptr(b), a.shape[0], float(grid_size), ptr(ws), ptr(out), current_stream()), "tulip_voxel_metrics")
    return out


def chamfer_distance(points1: torch.Tensor, points2: torch.Tensor):
    """-> (cd = mean(dist1) + mean(dist2) as a device scalar, dist1 [n1], dist2 [n2]); squared nearest-neighbour distances in float32
    (float64 DurLAR clouds are rounded to float32 first: the third-party extension the reference calls is a float kernel)."""
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
