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


# Ouster OS1-128 calibration constants of the DurLAR sensor as listed in evaluation.py:7-17 (beam elevation per row in degrees, the
# per-row column stagger, and the beam-origin offsets)
DURLAR_OFFSET_LUT = np.tile(np.array([48, 32, 16, 0]), 32)
DURLAR_ELEVATION_LUT = np.array([
    21.42, 21.12, 20.81, 20.5, 20.2, 19.9, 19.58, 19.26, 18.95, 18.65, 18.33, 18.02, 17.68, 17.37, 17.05, 16.73, 16.4, 16.08, 15.76, 15.43,
    15.1, 14.77, 14.45, 14.11, 13.78, 13.45, 13.13, 12.79, 12.44, 12.12, 11.77, 11.45, 11.1, 10.77, 10.43, 10.1, 9.74, 9.4, 9.06, 8.72,
    8.36, 8.02, 7.68, 7.34, 6.98, 6.63, 6.29, 5.95, 5.6, 5.25, 4.9, 4.55, 4.19, 3.85, 3.49, 3.15, 2.79, 2.44, 2.1, 1.75, 1.38, 1.03, 0.68,
    0.33, -0.03, -0.38, -0.73, -1.07, -1.45, -1.8, -2.14, -2.49, -2.85, -3.19, -3.54, -3.88, -4.26, -4.6, -4.95, -5.29, -5.66, -6.01,
    -6.34, -6.69, -7.05, -7.39, -7.73, -8.08, -8.44, -8.78, -9.12, -9.45, -9.82, -10.16, -10.5, -10.82, -11.19, -11.52, -11.85, -12.18,
    -12.54, -12.87, -13.2, -13.52, -13.88, -14.21, -14.53, -14.85, -15.2, -15.53, -15.84, -16.16, -16.5, -16.83, -17.14, -17.45, -17.8,
    -18.11, -18.42, -18.72, -19.06, -19.37, -19.68, -19.97, -20.31, -20.61, -20.92, -21.22])
DURLAR_ORIGIN_OFFSET = 0.015806
DURLAR_Z_OFFSET = 0.03618
DURLAR_ANGLE_OFF = np.pi * 4.2285 / 180.


def durlar_tables(rows, cols):
    """float64 tables of px_to_xyz (evaluation.py:27-45): cos / sin(encoder + azimuth), cos / sin(encoder) per column, cos / sin(elevation)
    per row, and the per-row column offsets of idx_from_px (:19-23)."""
    col = np.arange(cols)
    u = (cols + col) % cols
    encoder = 2.0 * np.pi - (u * (np.pi * 2.0 / cols))
    elevation = np.pi * DURLAR_ELEVATION_LUT[np.arange(rows)] / 180.
    return (np.cos(encoder + DURLAR_ANGLE_OFF), np.sin(encoder + DURLAR_ANGLE_OFF), np.cos(encoder), np.sin(encoder),
            np.cos(elevation), np.sin(elevation), DURLAR_OFFSET_LUT[:rows].astype(np.int32))


def range_to_points_durlar(img, maximum_range=120):
    """img (H, W) float32 -> (H*W, 3) float64 exactly as img_to_pcd_durlar (evaluation.py:47-58)."""
    img = np.asarray(img, np.float32)
    rows, cols = img.shape
    ca, sa, ce, se, cel, sel, off = durlar_tables(rows, cols)
    rr = (img * maximum_range) - np.float32(DURLAR_ORIGIN_OFFSET)          # float32, as numpy promotes `array - python float`
    x = rr * ca[None, :] * cel[:, None] + DURLAR_ORIGIN_OFFSET * ce[None, :]
    y = rr * sa[None, :] * cel[:, None] + DURLAR_ORIGIN_OFFSET * se[None, :]
    z = rr * sel[:, None]
    pts = np.stack((-x, -y, z + DURLAR_Z_OFFSET), axis=-1)
    col, row = np.arange(cols), np.arange(rows)
    idx = row[:, None] * cols + (col[None, :] + cols - off[:, None]) % cols
    out = np.zeros((rows * cols, 3))
    out[idx.reshape(-1)] = pts.reshape(-1, 3)
    return out


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
