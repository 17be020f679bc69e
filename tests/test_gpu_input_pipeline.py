"""Input pipeline kernel against the oracle that is pinned bit for bit to the reference's transform chains."""
import numpy as np
import pytest
import torch

from oracle import input_pipeline as P
from tests.test_oracle_input_pipeline import SIZES, raw_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset", ["kitti", "durlar", "carla"])
@pytest.mark.parametrize("log_transform", [True, False])
def test_preprocess_kernel_vs_oracle(dataset, log_transform):
    from tulip_b200.input_pipeline import preprocess
    raw = raw_frames()[dataset]
    out_size, in_size = SIZES[dataset]
    lo_o, hi_o = P.preprocess(raw, dataset, in_size[0], in_size[1], log_transform)
    lo, hi = preprocess(torch.from_numpy(raw).cuda(), dataset, in_size, log_transform)
    lo, hi = lo.cpu(), hi.cpu()
    if log_transform:      # log1pf vs torch.log1p: <= 1 ulp apart; zeros (filtered pixels) exact
        np.testing.assert_allclose(hi.numpy(), hi_o.numpy(), rtol=2e-7, atol=0)
        np.testing.assert_allclose(lo.numpy(), lo_o.numpy(), rtol=2e-7, atol=0)
        assert torch.equal(hi == 0, hi_o == 0)
    else:                  # scale + filter + row / column selection only: bit-exact
        assert torch.equal(hi, hi_o) and torch.equal(lo, lo_o)
    # the low-resolution input is exactly the selected rows / columns of the high-resolution target
    rf, cf = out_size[0] // in_size[0], out_size[1] // in_size[1]
    assert torch.equal(lo, hi[:, :, ::rf, ::cf])


def test_preprocess_single_channel_and_errors():
    from tulip_b200.input_pipeline import preprocess
    raw = torch.rand(3, 64, 1024, device="cuda") * 80
    lo, hi = preprocess(raw, "kitti", (16, 1024))
    assert torch.allclose(hi[:, 0], torch.log1p(raw * (1 / 80)), rtol=2e-7, atol=0) and tuple(lo.shape) == (3, 1, 16, 1024)
    with pytest.raises(ValueError):
        preprocess(raw, "kitti", (15, 1024))
    with pytest.raises(NotImplementedError):
        preprocess(raw, "nuscenes", (16, 1024))


def test_rimg_decode_kernel_bit_exact(tmp_path):
    """CARLA .rimg payloads -> fp32 frames (rimg_loader, datasets.py:181-193): bit-exact against the reference loader's output
    kept in the fixture and against the oracle on ragged sizes (tiles of 32 cut both ways), batched; then the carla chain."""
    import os
    from tulip_b200.input_pipeline import decode_rimg, preprocess, read_rimg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.npz"))
    file_a, file_b = g["rimg_a_file"].tobytes(), g["rimg_b_file"].tobytes()
    got = read_rimg([file_a])
    assert got.dtype == torch.float32 and np.array_equal(got[0].cpu().numpy(), g["rimg_a_frame"])
    path = tmp_path / "frame.rimg"
    path.write_bytes(file_b)
    frames = read_rimg([str(path), file_b])
    want, _ = P.rimg_decode(file_b)
    assert np.array_equal(frames[0].cpu().numpy(), want) and torch.equal(frames[0], frames[1])
    lo, hi = preprocess(frames[:, :, :, None].contiguous(), "carla", (16, 1024), True)
    lo_o, hi_o = P.preprocess(want[None, :, :, None], "carla", 16, 1024, True)
    np.testing.assert_allclose(hi[0].cpu().numpy(), hi_o[0].numpy(), rtol=2e-7, atol=0)
    np.testing.assert_allclose(lo[0].cpu().numpy(), lo_o[0].numpy(), rtol=2e-7, atol=0)
    rng = np.random.Generator(np.random.PCG64(5))
    for s0, s1, B in ((1, 1, 1), (33, 70, 3), (128, 2048, 2), (7, 31, 1)):
        rows = (rng.random((B, s1, s0), dtype=np.float32) * 100).astype(np.float16)
        out = decode_rimg(torch.from_numpy(rows).cuda().view(B, -1), (s0, s1)).cpu().numpy()
        for b in range(B):
            buf = np.array([s0, s1], dtype=np.uint64).tobytes() + rows[b].tobytes()
            assert np.array_equal(out[b], P.rimg_decode(buf)[0]), (s0, s1, b)
    with pytest.raises(ValueError):
        read_rimg([file_a, file_b])
    with pytest.raises(TypeError):
        decode_rimg(torch.zeros(4, device="cuda"), (2, 2))


def test_read_npy_feeds_preprocess(tmp_path):
    """`.npy` frames as the KITTI / DurLAR folders hold them ((H, W, 2) float32, npy_loader datasets.py:187-191) -> read_npy ->
    preprocess == the oracle chain on np.load(f)[..., 0]."""
    from tulip_b200.input_pipeline import preprocess, read_npy
    raw = raw_frames()["durlar"]
    paths = []
    for b in range(raw.shape[0]):
        paths.append(str(tmp_path / f"{b:08d}.npy"))
        np.save(paths[-1], raw[b])
    batch = read_npy(paths)
    assert batch.is_cuda and tuple(batch.shape) == raw.shape and torch.equal(batch.cpu(), torch.from_numpy(raw))
    out_size, in_size = SIZES["durlar"]
    lo, hi = preprocess(batch, "durlar", in_size, False)
    lo_o, hi_o = P.preprocess(np.stack([np.load(p)[..., 0].astype(np.float32) for p in paths]), "durlar", in_size[0], in_size[1], False)
    assert torch.equal(hi.cpu(), hi_o) and torch.equal(lo.cpu(), lo_o)
    np.save(str(tmp_path / "odd.npy"), raw[0, :8])
    with pytest.raises(ValueError):
        read_npy([paths[0], str(tmp_path / "odd.npy")])
