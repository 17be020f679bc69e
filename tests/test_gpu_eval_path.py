"""Evaluation path: the fused post-processing kernel against the numpy oracle, and the forward-only entry."""
import numpy as np
import pytest
import torch

from oracle.eval_post import CLIP_LO, eval_postprocess as oracle_post

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import tulip_b200.ops as o
    return o


@pytest.mark.parametrize("dataset,B,h,H,W", [("kitti", 1, 16, 64, 1024), ("kitti", 8, 16, 64, 1024), ("durlar", 3, 32, 128, 2048),
                                             ("carla", 2, 16, 64, 1024), ("kitti", 5, 4, 16, 36)])
@pytest.mark.parametrize("log_transform", [True, False])
def test_eval_postprocess_kernel_vs_oracle(ops, dataset, B, h, H, W, log_transform):
    g = torch.Generator().manual_seed(11)
    hi = torch.rand(B, 1, H, W, generator=g) * 0.7
    lo = hi[:, :, :: H // h, :].clone() + 0.01 * torch.rand(B, 1, h, W, generator=g)
    pred = hi + 0.05 * torch.randn(B, 1, H, W, generator=g)
    want, want_l = oracle_post(pred.numpy(), lo.numpy(), hi.numpy(), log_transform, dataset)
    out, losses = ops.eval_postprocess(pred.cuda(), lo.cuda(), hi.cuda(), log_transform, CLIP_LO[dataset], True)
    out, losses = out.cpu().numpy(), losses.cpu().numpy()
    # fp32 element-wise work: tolerance 1e-6 relative (expm1f vs numpy's expm1), exact zeros where the clip zeroes
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-7)
    assert ((out == 0) == (want == 0)).mean() >= 0.9999          # a value within 1 ulp of a clip bound may fall on the other side
    np.testing.assert_allclose(losses, want_l, rtol=1e-5, atol=1e-7)


def test_eval_postprocess_without_sensor_rows(ops):
    g = torch.Generator().manual_seed(12)
    hi = torch.rand(2, 1, 64, 1024, generator=g)
    lo = torch.rand(2, 1, 16, 512, generator=g)                   # carla with a narrower input: rows are not restored (:207-208)
    want, want_l = oracle_post(hi.numpy(), lo.numpy(), hi.numpy(), False, "carla")
    out, losses = ops.eval_postprocess(hi.cuda(), lo.cuda(), hi.cuda(), False, CLIP_LO["carla"], False)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    assert (losses.cpu().numpy()[:, 1] == 0).all()
    with pytest.raises(ValueError):
        ops.eval_postprocess(hi.cuda(), lo.cuda(), hi.cuda(), False, CLIP_LO["carla"], True)


def test_upsample_entry_matches_model_plus_oracle():
    """inference.upsample == model forward (eval) followed by the oracle post-processing; repeated calls replay the CUDA graph."""
    from oracle.params import TULIP_BASE, make_inputs, make_params
    from tests.test_gpu_model import build, load_params
    from tulip_b200.inference import upsample
    cfg = TULIP_BASE
    model = build(cfg).train()                                    # upsample() must switch to eval and restore the mode
    load_params(model, make_params(cfg, 5))
    model.cuda()
    lo, hi = make_inputs(cfg, 1, 6)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    outs = [upsample(model, lo_t, hi_t, "kitti") for _ in range(4)]
    assert model.training
    with torch.no_grad():
        pred, _, _ = model.eval()(lo_t, hi_t, eval=True)
    want, want_l = oracle_post(pred.cpu().numpy(), lo, hi, True, "kitti")
    for out, losses in outs:
        np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(losses.cpu().numpy(), want_l, rtol=1e-5, atol=1e-7)
    assert (outs[0][0][0, 0, ::4] == torch.expm1(lo_t[0, 0])).all()


def test_mc_dropout_aggregate_and_entry():
    """MCdrop()'s aggregation (engine_upsampling.py:423-427) as one kernel against the oracle, and the per-frame entry point."""
    import numpy as np
    import torch
    from oracle.eval_post import mc_dropout_aggregate
    from oracle.params import TULIP_BASE, make_inputs
    from tests.test_gpu_model import build
    from tulip_b200 import ops
    from tulip_b200.inference import mc_dropout_upsample
    g = torch.Generator().manual_seed(3)
    base = torch.rand(1, 1, 64, 1024, generator=g)
    preds = base + 0.02 * torch.randn(50, 1, 64, 1024, generator=g) * (torch.rand(1, 1, 64, 1024, generator=g) > 0.5)
    want, want_std = mc_dropout_aggregate(preds.numpy(), 0.03)
    got, got_std = ops.mc_dropout_aggregate(preds.cuda(), 0.03, return_std=True)
    np.testing.assert_allclose(got_std.cpu().numpy(), want_std, rtol=2e-5, atol=2e-6)
    keep = np.abs(want_std - 0.03 * preds.numpy().mean(0, keepdims=True)) > 1e-6
    np.testing.assert_allclose(got.cpu().numpy()[keep], want[keep], rtol=1e-6, atol=1e-7)
    model = build(TULIP_BASE).cuda().eval()
    lo, hi = make_inputs(TULIP_BASE, 1, 9)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    out = mc_dropout_upsample(model, lo_t, hi_t, iterations=18, iteration_batch=8)
    with torch.no_grad():
        single = model(lo_t, hi_t, mc_drop=True)
    # deterministic model (all dropout p = 0): every pass equals the single pass, std = 0, nothing is removed where mean > 0
    assert torch.allclose(out[single > 0], single[single > 0], rtol=1e-6, atol=1e-7)
