"""oracle/eval_metrics.py against the fixtures that oracle/make_golden_eval.py generated from the reference's own
util/evaluation.py functions (projections bit-exact, voxel metrics exact)."""
import hashlib
import os

import numpy as np

from oracle import eval_metrics as M

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.npz")


def images():
    rng = np.random.Generator(np.random.PCG64(123))                 # same stream as make_golden_eval.py
    gt = (rng.random((64, 1024), dtype=np.float32) * 0.6 + 0.03).astype(np.float32)
    pred = np.clip(gt + rng.normal(0, 0.004, gt.shape).astype(np.float32), 0, 1).astype(np.float32)
    pred[rng.random(gt.shape) < 0.05] = 0.0
    return gt, pred


def test_projection_and_metrics_match_reference_fixtures():
    g = np.load(GOLDEN)
    gt, pred = images()
    assert np.array_equal(gt[:, ::16], g["img_gt"]) and np.array_equal(pred[:, ::16], g["img_pred"])
    pk = M.range_to_points(pred, M.angle_tables_kitti(64, 1024), 80)
    assert np.array_equal(pk[:4096], g["kitti_points_head"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(pk.tobytes()).digest()[:8], np.uint8), g["kitti_points_sha"])
    pc = M.range_to_points(pred, M.angle_tables_carla(64, 1024), 80)
    assert np.array_equal(pc[:4096], g["carla_points_head"])
    p8, g8 = M.range_to_points(pred, M.angle_tables_kitti(64, 1024), 8), M.range_to_points(gt, M.angle_tables_kitti(64, 1024), 8)
    np.testing.assert_allclose(M.voxel_metrics(p8, g8, 0.1), g["voxel_metrics_range8_grid01"], rtol=0, atol=1e-15)
    cd, d1, d2 = M.chamfer_distance(g8[::16], p8[::16])
    assert abs(cd - g["chamfer_sub16_range8"][0]) <= 1e-7


def test_chamfer_of_identical_clouds_is_zero_and_symmetric():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(500, 3)).astype(np.float32)
    b = rng.normal(size=(300, 3)).astype(np.float32)
    assert M.chamfer_distance(a, a)[0] == 0
    assert abs(M.chamfer_distance(a, b)[0] - M.chamfer_distance(b, a)[0]) <= 1e-6
