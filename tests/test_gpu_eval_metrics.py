"""Evaluation metrics kernels against the oracle that is pinned to the reference's util/evaluation.py:
projections bit-exact, voxel metrics exact (integer set sizes), Chamfer nearest-neighbour distances bit-exact, means to 1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_metrics as M

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.npz")


def images():
    rng = np.random.Generator(np.random.PCG64(123))
    gt = (rng.random((64, 1024), dtype=np.float32) * 0.6 + 0.03).astype(np.float32)
    pred = np.clip(gt + rng.normal(0, 0.004, gt.shape).astype(np.float32), 0, 1).astype(np.float32)
    pred[rng.random(gt.shape) < 0.05] = 0.0
    return gt, pred


@pytest.mark.parametrize("dataset,rows", [("kitti", 64), ("kitti", 16), ("carla", 64)])
def test_range_to_points_bit_exact(dataset, rows):
    from tulip_b200 import metrics
    gt, pred = images()
    img = pred[:: 64 // rows]
    tables = M.angle_tables_kitti(rows, 1024) if dataset == "kitti" else M.angle_tables_carla(rows, 1024)
    want = M.range_to_points(img, tables, 80)
    got = metrics.range_to_points(torch.from_numpy(np.stack([img, img[::-1].copy()])).cuda(), dataset).cpu().numpy()
    assert np.array_equal(got[0], want)
    assert np.array_equal(got[1], M.range_to_points(img[::-1], tables, 80))
    if dataset == "kitti" and rows == 64:
        assert np.array_equal(got[0][:4096], np.load(GOLDEN)["kitti_points_head"])       # the reference's own output


@pytest.mark.parametrize("rng_range,grid", [(8, 0.1), (80, 0.1), (80, 0.5)])
def test_voxel_metrics_exact(rng_range, grid):
    from tulip_b200 import metrics
    gt, pred = images()
    t = M.angle_tables_kitti(64, 1024)
    pp, pg = M.range_to_points(pred, t, rng_range), M.range_to_points(gt, t, rng_range)
    want = M.voxel_metrics(pp, pg, grid)
    got = metrics.voxel_metrics(torch.from_numpy(pp).cuda(), torch.from_numpy(pg).cuda(), grid).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-15)
    if rng_range == 8 and grid == 0.1:
        np.testing.assert_allclose(got, np.load(GOLDEN)["voxel_metrics_range8_grid01"], rtol=0, atol=1e-15)   # reference dense grids


def test_chamfer_distance_vs_oracle():
    from tulip_b200 import metrics
    gt, pred = images()
    t = M.angle_tables_kitti(64, 1024)
    pp, pg = M.range_to_points(pred, t, 80)[::8], M.range_to_points(gt, t, 80)[::8]       # 8192 points each: seconds for numpy
    cd, d1, d2 = M.chamfer_distance(pg, pp)
    gcd, g1, g2 = metrics.chamfer_distance(torch.from_numpy(pg).cuda(), torch.from_numpy(pp[:8000].copy()).cuda())
    cd2, o1, o2 = M.chamfer_distance(pg, pp[:8000])
    assert np.array_equal(g1.cpu().numpy(), o1) and np.array_equal(g2.cpu().numpy(), o2)   # same expression order, no FMA
    assert abs(gcd.item() - cd2) <= 1e-6 * max(1.0, abs(cd2))


def test_evaluate_frame_full_size():
    """full KITTI frame (65536 points): properties -- identical images give zero distance and perfect overlap; the metric block
    of a perturbed frame agrees with the oracle's voxel metrics and is symmetric in the Chamfer term."""
    from tulip_b200 import metrics
    gt, pred = images()
    same = metrics.evaluate_frame(torch.from_numpy(gt).cuda(), torch.from_numpy(gt).cuda(), "kitti", 0.1)
    assert same["chamfer_dist"] == 0.0 and same["iou"] == 1.0 and same["f1"] == 1.0
    m = metrics.evaluate_frame(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), "kitti", 0.1)
    t = M.angle_tables_kitti(64, 1024)
    want = M.voxel_metrics(M.range_to_points(pred, t, 80), M.range_to_points(gt, t, 80), 0.1)
    np.testing.assert_allclose([m["iou"], m["precision"], m["recall"], m["f1"]], want, rtol=0, atol=1e-15)
    m2 = metrics.evaluate_frame(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(), "kitti", 0.1)
    assert abs(m["chamfer_dist"] - m2["chamfer_dist"]) <= 1e-6 and m["chamfer_dist"] > 0
    assert abs(m["precision"] - m2["recall"]) <= 1e-15


def images_durlar():
    rng = np.random.Generator(np.random.PCG64(123))                 # continue the stream of make_golden_eval.py past the kitti images
    gt = rng.random((64, 1024), dtype=np.float32); rng.normal(0, 0.004, gt.shape); rng.random(gt.shape)
    gd = (rng.random((128, 2048), dtype=np.float32) * 0.5 + 0.01).astype(np.float32)
    pd_ = np.clip(gd + rng.normal(0, 0.003, gd.shape).astype(np.float32), 0, 1).astype(np.float32)
    return gd, pd_


def test_durlar_projection_and_float64_voxel_metrics():
    from tulip_b200 import metrics
    gd, pd_ = images_durlar()
    g = np.load(GOLDEN)
    want = M.range_to_points_durlar(pd_, 120)
    got = metrics.range_to_points(torch.from_numpy(pd_).cuda(), "durlar")[0].cpu().numpy()
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert np.array_equal(got[:4096], g["durlar_points_head"])                              # the reference's own output
    pp = metrics.range_to_points(torch.from_numpy(pd_).cuda(), "durlar", 12)[0]
    pg = metrics.range_to_points(torch.from_numpy(gd).cuda(), "durlar", 12)[0]
    vm = metrics.voxel_metrics(pp, pg, 0.1).cpu().numpy()
    np.testing.assert_allclose(vm, g["durlar_voxel_metrics_range12_grid01"], rtol=0, atol=1e-15)   # reference dense grids, float64
    m = metrics.evaluate_frame(torch.from_numpy(pd_).cuda(), torch.from_numpy(gd).cuda(), "durlar", 0.1)
    assert m["chamfer_dist"] > 0 and 0 < m["iou"] < 1
