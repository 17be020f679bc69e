"""Oracle index ops vs the known answers generated from the unmodified reference
(tests/golden/index_ops.npz, written by oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import index_ops as I
from oracle.params import TULIP_BASE
from oracle.tulip_oracle import drop_path_rates


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "index_ops.npz"))


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_relative_position_index(g):
    idx = I.relative_position_index((2, 8))
    assert idx.dtype == np.int64 and np.array_equal(idx, g["rel_index_2x8"])
    assert sha16(idx) == "4ddf27b1d7ef65c1"
    assert idx[0].tolist() == [22, 21, 20, 19, 18, 17, 16, 15, 7, 6, 5, 4, 3, 2, 1, 0]


def test_window_partition_and_reverse(g):
    x = np.arange(16 * 256, dtype=np.float32).reshape(1, 16, 256, 1)
    w = I.window_partition(x, (2, 8))
    assert np.array_equal(w, g["partition_16x256"]) and sha16(w) == "18cb13e7e85d8ed1"
    assert np.array_equal(I.window_reverse(w, (2, 8), 16, 256), x)


@pytest.mark.parametrize("key,H,W,win,shift", [
    ("mask_16x256", 16, 256, (2, 8), (1, 4)),
    ("mask_32x512", 32, 512, (2, 8), (1, 4)),
    ("mask_backup_1x16", 1, 16, (1, 16), (0, 8)),
    ("mask_backup_1x64", 1, 64, (1, 16), (0, 8)),
])
def test_shift_mask(g, key, H, W, win, shift):
    a = I.shift_mask_slices(H, W, win, shift)
    b = I.shift_mask_closed_form(H, W, win, shift)
    assert a.dtype == np.float32 and np.array_equal(a, g[key]) and np.array_equal(b, g[key])


def test_shift_mask_counts(g):
    m = g["mask_16x256"]
    assert sha16(m) == "e85f862a06f42493" and int((m == -100).sum()) == 5056
    assert int((g["mask_backup_1x16"] == -100).sum()) == 128


@pytest.mark.parametrize("H,W", [(2, 8), (4, 64), (8, 128), (2, 32), (16, 512)])
def test_shift_mask_closed_form_general(H, W):
    assert np.array_equal(I.shift_mask_slices(H, W, (2, 8), (1, 4)), I.shift_mask_closed_form(H, W, (2, 8), (1, 4)))


def test_merge_shuffle_pad_roll(g):
    assert np.array_equal(I.merge_2x2(np.arange(16.).reshape(1, 4, 4, 1)), g["merge_4x4"])
    assert np.array_equal(I.pixel_shuffle_nchw(np.arange(8.).reshape(1, 8, 1, 1), 2), g["pixel_shuffle_r2"])
    x = np.arange(2 * 32 * 2 * 3, dtype=np.float32).reshape(2, 32, 2, 3)
    assert np.array_equal(I.pixel_shuffle_nchw(x, 4), g["pixel_shuffle_r4"])
    nhwc = I.pixel_shuffle_nhwc(x.transpose(0, 2, 3, 1), 4)
    assert np.array_equal(nhwc.transpose(0, 3, 1, 2), g["pixel_shuffle_r4"])
    assert np.array_equal(I.circular_pad_w(np.arange(1024.).reshape(1, 1, 1, 1024)), g["circ_pad_1024"])
    assert np.array_equal(I.cyclic_shift(np.arange(64.).reshape(1, 4, 16, 1), -1, -4), g["roll_m1_m4"])


def test_config_known_answers(g):
    assert TULIP_BASE.upscale_factor == int(g["upscale_factor_kitti"]) == 4
    assert tuple(g["grid_kitti"]) == TULIP_BASE.grid == (16, 256)
    enc, dec = drop_path_rates(TULIP_BASE)
    assert np.allclose(np.array(enc), g["drop_rates_enc"], atol=1e-7)
    assert np.allclose(np.array(dec), g["drop_rates_dec"], atol=1e-7)
