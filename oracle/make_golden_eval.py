"""Pin oracle/eval_metrics.py against the UNMODIFIED tulip/util/evaluation.py and write tests/golden/eval_metrics.npz.
Build container only (needs /root/reference):   python -m oracle.make_golden_eval"""
import os
import sys
import types

import numpy as np

from . import eval_metrics as M

REF_ROOT = os.environ.get("TULIP_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference_evaluation():
    if "chamfer_distance" not in sys.modules:                    # un-vendored CUDA extension, only used by chamfer_distance()
        cd = types.ModuleType("chamfer_distance")
        cd.ChamferDistance = object
        sys.modules["chamfer_distance"] = cd
    sys.path.insert(0, os.path.join(REF_ROOT, "tulip"))
    import util.evaluation as E
    return E


def main():
    E = import_reference_evaluation()
    rng = np.random.Generator(np.random.PCG64(123))
    gt = (rng.random((64, 1024), dtype=np.float32) * 0.6 + 0.03).astype(np.float32)
    pred = np.clip(gt + rng.normal(0, 0.004, gt.shape).astype(np.float32), 0, 1).astype(np.float32)
    pred[rng.random(gt.shape) < 0.05] = 0.0                      # pixels the range clip zeroed
    out = {"img_gt": gt[:, ::16].copy(), "img_pred": pred[:, ::16].copy()}
    # projections: the reference's own functions vs the oracle, bit for bit
    ref_k = E.img_to_pcd_kitti(pred, maximum_range=80)
    ora_k = M.range_to_points(pred, M.angle_tables_kitti(64, 1024), 80)
    assert ref_k.dtype == np.float32 and np.array_equal(ref_k, ora_k), "kitti projection differs from the reference"
    ref_c = E.img_to_pcd_carla(pred, maximum_range=80)
    ora_c = M.range_to_points(pred, M.angle_tables_carla(64, 1024), 80)
    assert np.array_equal(ref_c.astype(np.float32), ora_c), "carla projection differs from the reference"
    ref_k16 = E.img_to_pcd_kitti(pred[::4], maximum_range=80, low_res=True)
    assert np.array_equal(ref_k16, M.range_to_points(pred[::4], M.angle_tables_kitti(16, 1024), 80))
    # voxel metrics through the reference's dense grids (small extent so the boolean grids stay small) vs the oracle's sets
    pk, gk = E.img_to_pcd_kitti(pred, 8), E.img_to_pcd_kitti(gt, 8)
    pcd_all = np.vstack((pk, gk))
    mn, mx = np.min(pcd_all, axis=0), np.max(pcd_all, axis=0)
    vp, vg = E.voxelize_point_cloud(pk, 0.1, mn, mx), E.voxelize_point_cloud(gk, 0.1, mn, mx)
    iou, prec, rec = E.calculate_metrics(vp, vg)
    ref_m = np.array([iou, prec, rec, 2 * prec * rec / (prec + rec)])
    ora_m = M.voxel_metrics(pk, gk, 0.1)
    assert np.allclose(ref_m, ora_m, rtol=0, atol=1e-15), (ref_m, ora_m)
    # DurLAR (Ouster LUT projection, float64) and the float64 voxel metrics on it
    gd = (rng.random((128, 2048), dtype=np.float32) * 0.5 + 0.01).astype(np.float32)
    pd_ = np.clip(gd + rng.normal(0, 0.003, gd.shape).astype(np.float32), 0, 1).astype(np.float32)
    ref_d = E.img_to_pcd_durlar(pd_, maximum_range=120)
    assert ref_d.dtype == np.float64 and np.array_equal(ref_d, M.range_to_points_durlar(pd_, 120)), "durlar projection differs"
    assert np.array_equal(E.offset_lut, M.DURLAR_OFFSET_LUT) and np.array_equal(E.elevation_lut, M.DURLAR_ELEVATION_LUT)
    pdd, gdd = E.img_to_pcd_durlar(pd_, 12), E.img_to_pcd_durlar(gd, 12)
    alld = np.vstack((pdd, gdd))
    mnd, mxd = np.min(alld, axis=0), np.max(alld, axis=0)
    ioud, precd, recd = E.calculate_metrics(E.voxelize_point_cloud(pdd, 0.1, mnd, mxd), E.voxelize_point_cloud(gdd, 0.1, mnd, mxd))
    ref_md = np.array([ioud, precd, recd, 2 * precd * recd / (precd + recd)])
    assert np.allclose(ref_md, M.voxel_metrics(pdd, gdd, 0.1), rtol=0, atol=1e-15)
    out.update({"durlar_points_head": ref_d[:4096].copy(), "durlar_voxel_metrics_range12_grid01": ref_md,
                "durlar_points_sha": np.frombuffer(__import__("hashlib").sha256(ref_d.tobytes()).digest()[:8], np.uint8)})
    sub = slice(0, 65536, 16)
    cd, d1, d2 = M.chamfer_distance(gk[sub], pk[sub])
    out.update({"kitti_points_sha": np.frombuffer(__import__("hashlib").sha256(ref_k.tobytes()).digest()[:8], np.uint8),
                "kitti_points_head": ref_k[:4096].copy(), "carla_points_head": ref_c[:4096].astype(np.float32),
                "voxel_metrics_range8_grid01": ref_m, "chamfer_sub16_range8": np.array([cd], np.float32)})
    # the full-size images the fixtures were computed from are regenerated from the seed by the tests
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "eval_metrics.npz"), **out)
    print("eval metrics oracle pinned: kitti/carla projections bit-exact, voxel metrics", ref_m, "chamfer (oracle)", cd)


if __name__ == "__main__":
    main()
