"""Bit-exact parity of the index arithmetic (window partition / cyclic shift / mask / bias gather / merge gather /
pixel shuffle) with the oracle's numpy restatement and the reference-generated known answers.  Called through
the C ABI; these use the same device functions as the fused attention / LayerNorm / GEMM kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import index_ops as I

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from tulip_b200 import ops
    return ops


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "index_ops.npz"))


def rnd_bf16(*shape, seed=0):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=gen).to(torch.bfloat16)


@pytest.mark.parametrize("B,H,W,C,win,shift", [
    (2, 16, 256, 96, (2, 8), (0, 0)), (2, 16, 256, 96, (2, 8), (1, 4)), (3, 2, 32, 768, (2, 8), (1, 4)),
    (1, 1, 16, 1536, (1, 16), (0, 8)), (2, 1, 64, 96, (1, 16), (0, 0)), (1, 32, 512, 96, (2, 8), (1, 4)),
])
def test_window_partition_reverse_roll(ops, B, H, W, C, win, shift):
    x = rnd_bf16(B, H, W, C)
    want = I.window_partition(I.cyclic_shift(x.float().numpy(), -shift[0], -shift[1]), win)
    got = ops.window_partition(x.cuda(), win, shift)
    assert torch.equal(got.float().cpu(), torch.from_numpy(want))
    back = ops.window_reverse(got, B, H, W, win, shift)          # reverse + roll(+sh,+sw) (tulip.py:320,323)
    assert torch.equal(back.cpu(), x)


def test_window_partition_known_answer(ops, g):
    x = torch.arange(16 * 256, dtype=torch.float32).view(1, 16, 256, 1).expand(1, 16, 256, 8).contiguous()
    got = ops.window_partition((x % 256).cuda(), (2, 8), (0, 0))[..., :1].float().cpu().numpy()   # bf16 holds ints < 257 exactly
    assert np.array_equal(got, g["partition_16x256"] % 256)


@pytest.mark.parametrize("key,H,W,win,shift", [
    ("mask_16x256", 16, 256, (2, 8), (1, 4)), ("mask_32x512", 32, 512, (2, 8), (1, 4)),
    ("mask_backup_1x16", 1, 16, (1, 16), (0, 8)), ("mask_backup_1x64", 1, 64, (1, 16), (0, 8)),
])
def test_shift_mask(ops, g, key, H, W, win, shift):
    got = ops.shift_mask(H, W, win, shift).cpu().numpy()
    assert np.array_equal(got, g[key]) and np.array_equal(got, I.shift_mask_slices(H, W, win, shift))


def test_rel_bias_gather(ops, g):
    table = torch.randn(45, 6)
    got = ops.rel_bias_gather(table.cuda(), (2, 8)).cpu()
    idx = torch.from_numpy(g["rel_index_2x8"])
    want = table[idx.view(-1)].view(16, 16, -1).permute(2, 0, 1)
    assert torch.equal(got, want)


def test_merge_gather(ops, g):
    x = rnd_bf16(2, 8, 16, 96)
    assert torch.equal(ops.merge_gather(x.cuda()).float().cpu(), torch.from_numpy(I.merge_2x2(x.float().numpy())))
    small = torch.arange(16.).view(1, 4, 4, 1).expand(1, 4, 4, 8).contiguous()
    assert np.array_equal(ops.merge_gather(small.cuda()).float().cpu().numpy()[..., ::8], g["merge_4x4"])


@pytest.mark.parametrize("r,Cout", [(2, 48), (4, 96)])
def test_pixel_shuffle(ops, r, Cout):
    x = rnd_bf16(2, 4, 8, Cout * r * r)
    want = I.pixel_shuffle_nhwc(x.float().numpy(), r)
    assert torch.equal(ops.pixel_shuffle(x.cuda(), r).float().cpu(), torch.from_numpy(want))
