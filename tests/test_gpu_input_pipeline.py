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
