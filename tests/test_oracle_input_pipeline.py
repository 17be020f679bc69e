"""oracle/input_pipeline.py against the fixture that oracle/make_golden_input.py wrote from the reference's own transform classes."""
import os

import numpy as np

from oracle import input_pipeline as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.npz")
SIZES = {"kitti": ((64, 1024), (16, 1024)), "durlar": ((128, 2048), (32, 2048)), "carla": ((64, 1024), (16, 512))}


def raw_frames():
    rng = np.random.Generator(np.random.PCG64(77))                  # same stream as make_golden_input.py
    out = {}
    for dataset, (out_size, _) in SIZES.items():
        raw = (rng.random((2, *out_size, 2), dtype=np.float32) * 130.0).astype(np.float32)
        raw[rng.random(raw.shape) < 0.03] = 0.0
        out[dataset] = raw
    return out


def test_input_pipeline_oracle_matches_reference_fixture():
    g = np.load(GOLDEN)
    for dataset, raw in raw_frames().items():
        out_size, in_size = SIZES[dataset]
        assert np.array_equal(raw[:, ::8, ::16], g[f"{dataset}_raw"])
        lo, hi = P.preprocess(raw, dataset, in_size[0], in_size[1], True)
        assert tuple(lo.shape) == (2, 1, *in_size) and tuple(hi.shape) == (2, 1, *out_size)
        assert np.array_equal(hi.numpy()[:, :, ::8, ::16], g[f"{dataset}_hi"])
        assert abs(float(lo.double().sum()) - g[f"{dataset}_lo_sum"][0]) == 0
        if dataset != "kitti":
            assert (hi == 0).any()                                  # the range filter removed something


def test_rimg_decode_oracle_matches_reference_fixture():
    """oracle rimg_decode against what the reference's rimg_loader (datasets.py:181-193) returned for the same file bytes."""
    g = np.load(GOLDEN)
    frame, rows = P.rimg_decode(g["rimg_a_file"].tobytes())
    assert frame.dtype == np.float32 and np.array_equal(frame, g["rimg_a_frame"])
    assert rows.shape == frame.shape[::-1]
    assert frame[0, 0] == np.float32(rows[-1, -1]) and frame[-1, 0] == np.float32(rows[-1, 0])       # both axes flipped
    frame_b, _ = P.rimg_decode(g["rimg_b_file"].tobytes())
    assert frame_b.shape == tuple(g["rimg_b_seed"])
    assert [float(frame_b.astype(np.float64).sum()), float(frame_b[3, 17]), float(frame_b[-1, 0])] == list(g["rimg_b_frame_sum"])
    lo, hi = P.preprocess(frame_b[None, :, :, None], "carla", 16, 1024, True)
    assert [float(hi.double().sum()), float(lo.double().sum())] == list(g["rimg_b_hi_sum"])
