"""Golden fixtures of the head / upsampling variants no shipped script selects, generated from the UNMODIFIED reference
(build container only):

    python -m oracle.make_golden_r3

  tests/golden/model_expanding_kitti_b1.npz   tulip_base with patch_unmerging=False: PatchExpanding (tulip.py:126-141) in the
                                              decoder and as first_patch_expanding (:565), PixelShuffleHead kept
  tests/golden/model_expanding_head_kitti_b1.npz   tulip_base with patch_unmerging=False AND pixel_shuffle=False:
                                              FinalPatchExpanding head as well (tulip.py:144-159, 582, 727-729)

Both: batch 1, forward + backward, oracle checked against the reference on the same PCG64 parameters before writing.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Earlier fixtures are not touched."""
from __future__ import annotations

import dataclasses
import os

import torch

import oracle.make_golden as MG
from .make_golden import import_reference, model_fixture
from .params import TULIP_BASE

EXPANDING = dataclasses.replace(TULIP_BASE, patch_unmerging=False)
EXPANDING_HEAD = dataclasses.replace(TULIP_BASE, patch_unmerging=False, pixel_shuffle=False)


def build_variant(T, cfg, large):
    fn = T.tulip_large if large else T.tulip_base
    return fn(img_size=tuple(cfg.img_size), target_img_size=tuple(cfg.target_img_size), patch_size=tuple(cfg.patch_size),
              in_chans=cfg.in_chans, window_size=list(cfg.window_size), swin_v2=False, pixel_shuffle=cfg.pixel_shuffle,
              circular_padding=True, log_transform=cfg.log_transform, patch_unmerging=cfg.patch_unmerging)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    T = import_reference()
    saved = MG.build_reference
    MG.build_reference = build_variant
    try:
        model_fixture(T, "model_expanding_kitti_b1", EXPANDING, False, batch=1, pseed=8, xseed=9, store_pred_stride=4)
        model_fixture(T, "model_expanding_head_kitti_b1", EXPANDING_HEAD, False, batch=1, pseed=10, xseed=11, store_pred_stride=4)
    finally:
        MG.build_reference = saved


if __name__ == "__main__":
    main()
