"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's range-image -> point-cloud projection and voxel / Chamfer metrics
(tulip/util/evaluation.py:52-116, :125-175; used by engine_upsampling.py:223-276).  Only tests/, smoke() and bench.py's cpu_baseline
may import this module.

Pinned: `python -m oracle.make_golden_eval` (build container only) imports the UNMODIFIED tulip/util/evaluation.py (with an import
stub for the un-vendored `chamfer_distance` package) and checks img_to_pcd_kitti / img_to_pcd_carla / voxelize_point_cloud +
calculate_metrics against this file, writing tests/golden/eval_metrics.npz.
Chamfer distance: the reference calls the third-party CUDA extension github.com/otaheri/chamfer_distance (README.md:22-24, no
version pinned, not vendored).  Its published algorithm -- for every point the SQUARED Euclidean distance to its nearest neighbour
in the other cloud, both directions -- is restated here; evaluation.py:125-134 then adds the two means.  Parity for this one function
is therefore anchored on the reference's call site, not on the extension's bits ("parity unpinned" for the extension itself)."""
import numpy as np


def angle_tables_kitti(image_rows=64, image_cols=1024):
    """sin/cos tables of evaluation.py:52-72 (float32 arithmetic exactly as numpy performs it there)."""
    ang_start_y = 24.8
    ang_res_y = 26.8 / (image_rows - 1)
    ang_res_x = 360 / image_cols
    rows = np.arange(image_rows, dtype=np.float64)
    cols = np.arange(image_cols, dtype=np.float64)
    vertical = np.float32(rows * ang_res_y) - ang_start_y
    horizon = -np.float32(cols + 1 - (image_cols / 2)) * ang_res_x + 90.0
    vertical = vertical / 180.0 * np.pi
    horizon = horizon / 180.0 * np.pi
    return (np.sin(horizon).astype(np.float32), np.cos(horizon).astype(np.float32),
            np.sin(vertical).astype(np.float32), np.cos(vertical).astype(np.float32))


def angle_tables_carla(rows, cols):
    """evaluation.py:90-104."""
    v = np.deg2rad(np.linspace(start=-15, stop=15, num=rows).astype(np.float32))
    h = np.deg2rad(np.linspace(start=-180, stop=180, num=cols, endpoint=False).astype(np.float32))
    return np.sin(h), np.cos(h), np.sin(v), np.cos(v)


def range_to_points(img, tables, maximum_range):
    """img (H,W) float32 normalised range -> (H*W, 3) float32, row-major over (row, col): evaluation.py:75-84 / :106-114."""
    sin_h, cos_h, sin_v, cos_v = tables
    r = np.asarray(img, np.float32) * np.float32(maximum_range) if not isinstance(maximum_range, int) else np.asarray(img, np.float32) * maximum_range
    r = r.astype(np.float32)
    x = (sin_h[None, :] * cos_v[:, None]) * r
    y = (cos_h[None, :] * cos_v[:, None]) * r
    z = sin_v[:, None] * r
    return np.stack((x, y, z), axis=-1).reshape(-1, 3).astype(np.float32)


def voxel_metrics(pcd_pred, pcd_gt, grid_size):
    """engine_upsampling.py:259-276 + evaluation.py:148-175 with sets of voxel indices instead of dense boolean grids
    (identical counts; the dense grids of a KITTI frame at grid 0.1 are ~1600 x 1600 x 300 booleans each)."""
    pcd_all = np.vstack((pcd_pred, pcd_gt))
    min_coord = np.min(pcd_all, axis=0)
    ip = ((pcd_pred - min_coord) / grid_size).astype(int)
    ig = ((pcd_gt - min_coord) / grid_size).astype(int)
    sp = set(map(tuple, ip.tolist()))
    sg = set(map(tuple, ig.tolist()))
    inter = len(sp & sg)
    union = len(sp | sg)
    iou = inter / union
    precision = inter / len(sp)
    recall = inter / len(sg)
    f1 = 2 * (precision * recall) / (precision + recall)
    return np.array([iou, precision, recall, f1], np.float64)


def chamfer_distance(points1, points2, chunk=2048):
    """evaluation.py:125-134: mean over points1 of the squared distance to the nearest point2, plus the mirrored term."""
    a = np.asarray(points1, np.float32)
    b = np.asarray(points2, np.float32)

    def one_way(p, q):
        out = np.empty(len(p), np.float32)
        for i in range(0, len(p), chunk):
            d = p[i:i + chunk, None, :] - q[None, :, :]
            out[i:i + chunk] = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).min(axis=1)
        return out
    d1, d2 = one_way(a, b), one_way(b, a)
    return np.float32(d1.mean(dtype=np.float32) + d2.mean(dtype=np.float32)), d1, d2
