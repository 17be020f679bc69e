"""numpy restatement of the integer / index operations on the TULIP hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates (paths relative to /root/reference).  These ops are
pure data movement or integer logic, so parity with the CUDA path is bit-exact.
"""
from __future__ import annotations

import numpy as np


def window_partition(x: np.ndarray, win: tuple[int, int]) -> np.ndarray:
    """(B,H,W,C) -> (B*H/Mh*W/Mw, Mh, Mw, C), windows ordered (B, Nh, Nw).

    Reference: tulip/model/tulip.py:248-252 (einops
    'B (Nh Mh) (Nw Mw) C -> (B Nh Nw) Mh Mw C').
    """
    B, H, W, C = x.shape
    Mh, Mw = win
    assert H % Mh == 0 and W % Mw == 0
    x = x.reshape(B, H // Mh, Mh, W // Mw, Mw, C)
    return np.ascontiguousarray(x.transpose(0, 1, 3, 2, 4, 5)).reshape(-1, Mh, Mw, C)


def window_reverse(xw: np.ndarray, win: tuple[int, int], H: int, W: int) -> np.ndarray:
    """Inverse of window_partition. Reference: tulip/model/tulip.py:320."""
    Mh, Mw = win
    Nh, Nw = H // Mh, W // Mw
    C = xw.shape[-1]
    B = xw.shape[0] // (Nh * Nw)
    x = xw.reshape(B, Nh, Nw, Mh, Mw, C).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(x).reshape(B, H, W, C)


def cyclic_shift(x: np.ndarray, sh: int, sw: int) -> np.ndarray:
    """torch.roll(x, (sh, sw), dims=(1, 2)): out[b,h,w] = x[b,(h-sh)%H,(w-sw)%W].

    Reference: tulip/model/tulip.py:290 (negative shifts) and :323 (positive).
    """
    return np.roll(x, shift=(sh, sw), axis=(1, 2))


def relative_position_index(win: tuple[int, int]) -> np.ndarray:
    """(Mh*Mw, Mh*Mw) int64 table into the (2Mh-1)(2Mw-1)-row bias table.

    Reference: tulip/model/tulip.py:228-240.
    idx[i,j] = (ri-rj+Mh-1)*(2Mw-1) + (ci-cj+Mw-1) with i=(ri,ci) row-major.
    """
    Mh, Mw = win
    r = np.arange(Mh * Mw) // Mw
    c = np.arange(Mh * Mw) % Mw
    dr = r[:, None] - r[None, :] + (Mh - 1)
    dc = c[:, None] - c[None, :] + (Mw - 1)
    return (dr * (2 * Mw - 1) + dc).astype(np.int64)


def shift_mask_slices(H: int, W: int, win: tuple[int, int], shift: tuple[int, int]) -> np.ndarray:
    """(nW, L, L) float32 in {0,-100}: restates the reference's slice-fill procedure.

    Reference: tulip/model/tulip.py:254-280 (python slice semantics kept,
    including slice(-0, None) == everything when a shift component is 0).
    """
    Mh, Mw = win
    sh, sw = shift
    assert H % Mh == 0 and W % Mw == 0, "H or W is not divisible by window_size"
    img = np.zeros((1, H, W, 1), dtype=np.float32)
    h_slices = (slice(0, -Mh), slice(-Mh, -sh), slice(-sh, None))
    w_slices = (slice(0, -Mw), slice(-Mw, -sw), slice(-sw, None))
    cnt = 0
    for hs in h_slices:
        for ws in w_slices:
            img[:, hs, ws, :] = cnt
            cnt += 1
    mw = window_partition(img, win).reshape(-1, Mh * Mw)
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, np.float32(-100.0), np.float32(0.0)).astype(np.float32)


def shift_mask_closed_form(H: int, W: int, win: tuple[int, int], shift: tuple[int, int]) -> np.ndarray:
    """Closed form of shift_mask_slices used by the CUDA kernels.

    Region id on the rolled grid: hreg(h) = [h >= H-Mh] + [h >= H-sh] (second
    term only when sh > 0; with sh == 0 every row falls in the last slice so the
    row component is constant), likewise for w; id = 3*hreg + wreg;
    mask[win,i,j] = -100 if id_i != id_j else 0.
    """
    Mh, Mw = win
    sh, sw = shift
    h = np.arange(H)
    w = np.arange(W)
    if sh > 0:
        hreg = (h >= H - Mh).astype(np.int64) + (h >= H - sh).astype(np.int64)
    else:
        hreg = np.zeros_like(h)
    if sw > 0:
        wreg = (w >= W - Mw).astype(np.int64) + (w >= W - sw).astype(np.int64)
    else:
        wreg = np.zeros_like(w)
    rid = (3 * hreg[:, None] + wreg[None, :]).astype(np.float32).reshape(1, H, W, 1)
    mw = window_partition(rid, win).reshape(-1, Mh * Mw)
    return np.where(mw[:, None, :] != mw[:, :, None], np.float32(-100.0), np.float32(0.0)).astype(np.float32)


def merge_2x2(x: np.ndarray) -> np.ndarray:
    """(B,H,W,C) -> (B,H/2,W/2,4C), channel blocks [(0,0),(1,0),(0,1),(1,1)].

    Reference: tulip/model/tulip.py:92-99 (PatchMerging.merging).
    """
    return np.concatenate([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], axis=-1)


def pixel_shuffle_nchw(x: np.ndarray, r: int) -> np.ndarray:
    """nn.PixelShuffle(r): out[b,c,h*r+i,w*r+j] = in[b,c*r*r+i*r+j,h,w].

    Reference: tulip/model/tulip.py:115,120 and :171,176 (torch.nn.PixelShuffle).
    """
    B, Crr, H, W = x.shape
    C = Crr // (r * r)
    x = x.reshape(B, C, r, r, H, W).transpose(0, 1, 4, 2, 5, 3)
    return np.ascontiguousarray(x).reshape(B, C, H * r, W * r)


def pixel_shuffle_nhwc(x: np.ndarray, r: int) -> np.ndarray:
    """NHWC view of the same op: out[b,h*r+i,w*r+j,c] = in[b,h,w,c*r*r+i*r+j].

    This is what PatchUnmerging computes between its two permutes
    (tulip/model/tulip.py:117-123).
    """
    B, H, W, Crr = x.shape
    C = Crr // (r * r)
    x = x.reshape(B, H, W, C, r, r).transpose(0, 1, 4, 2, 5, 3)
    return np.ascontiguousarray(x).reshape(B, H * r, W * r, C)


def circular_pad_w(x: np.ndarray, left: int = 2, right: int = 2) -> np.ndarray:
    """F.pad(x, (2,2,0,0), 'circular') on the last axis. Reference: tulip.py:59-61."""
    return np.concatenate([x[..., -left:], x, x[..., :right]], axis=-1)
