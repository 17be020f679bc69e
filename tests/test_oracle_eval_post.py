"""Pins oracle/eval_post.py against the reference statements of evaluate() (engine_upsampling.py:174-244), executed one by
one in torch on the same inputs (B = 1, exactly as the reference's evaluation loader feeds them)."""
import numpy as np
import pytest
import torch

from oracle.eval_post import eval_postprocess


def reference_statements(pred_img, images_low_res, images_high_res, dataset, log_transform, h_low_res, h_high_res):
    downsampling_factor = h_high_res // h_low_res                                  # engine_upsampling.py:134
    if log_transform:                                                              # :177-180
        pred_img = torch.expm1(pred_img)
        images_high_res = torch.expm1(images_high_res)
        images_low_res = torch.expm1(images_low_res)
    if dataset in ("carla", "kitti"):                                              # :183-188
        pred_img = torch.where((pred_img >= 2 / 80) & (pred_img <= 1), pred_img, 0)
    elif dataset == "durlar":
        pred_img = torch.where((pred_img >= 0.3 / 120) & (pred_img <= 1), pred_img, 0)
    loss_map = (pred_img - images_high_res).abs()                                  # :192-193
    pixel_loss_one_input = loss_map.mean()
    images_low_res = images_low_res.permute(0, 2, 3, 1).squeeze().numpy()          # :195-203
    pred_img = pred_img.permute(0, 2, 3, 1).squeeze().numpy().copy()
    low_res_index = range(0, h_high_res, downsampling_factor)                      # :214 / :237
    pred_low_res_part = pred_img[low_res_index, :]
    loss_low_res_part = np.abs(pred_low_res_part - images_low_res).mean()          # :216-219
    pred_img[low_res_index, :] = images_low_res                                    # :221
    return pred_img, float(pixel_loss_one_input), float(loss_low_res_part)


@pytest.mark.parametrize("dataset,h,H,W", [("kitti", 16, 64, 1024), ("durlar", 32, 128, 2048), ("carla", 16, 64, 1024)])
@pytest.mark.parametrize("log_transform", [True, False])
def test_eval_postprocess_oracle_matches_reference_statements(dataset, h, H, W, log_transform):
    g = torch.Generator().manual_seed(7)
    hi = torch.rand(1, 1, H, W, generator=g) * 0.7
    lo = hi[:, :, :: H // h, :].clone()
    pred = hi + 0.05 * torch.randn(1, 1, H, W, generator=g)                        # some values leave [clip_lo, 1] and get zeroed
    want_img, want_pix, want_low = reference_statements(pred.clone(), lo.clone(), hi.clone(), dataset, log_transform, h, H)
    out, losses = eval_postprocess(pred.numpy(), lo.numpy(), hi.numpy(), log_transform, dataset)
    np.testing.assert_allclose(out[0, 0], want_img, rtol=1e-6, atol=1e-7)
    assert abs(losses[0, 0] - want_pix) <= 1e-6 * max(1.0, abs(want_pix))
    assert abs(losses[0, 1] - want_low) <= 1e-6 * max(1.0, abs(want_low))
    assert (out[0, 0] == 0).any() and (out[0, 0, :: H // h] == (np.expm1(lo.numpy()[0, 0]) if log_transform else lo.numpy()[0, 0])).all()


def test_eval_postprocess_carla_width_mismatch_keeps_prediction_rows():
    g = torch.Generator().manual_seed(9)
    hi = torch.rand(2, 1, 64, 1024, generator=g)
    lo = torch.rand(2, 1, 16, 512, generator=g)
    out, losses = eval_postprocess(hi.numpy(), lo.numpy(), hi.numpy(), False, "carla")   # engine_upsampling.py:207-208
    assert (losses[:, 1] == 0).all() and losses.shape == (2, 2)


def test_mc_dropout_aggregate_oracle_matches_reference_statements():
    """engine_upsampling.py:423-427 executed as written, against the numpy restatement."""
    from oracle.eval_post import mc_dropout_aggregate
    g = torch.Generator().manual_seed(3)
    base = torch.rand(1, 1, 16, 64, generator=g)
    pred_img_iteration = base + 0.02 * torch.randn(50, 1, 16, 64, generator=g) * (torch.rand(1, 1, 16, 64, generator=g) > 0.5)
    noise_threshold = 0.03
    pred_img = torch.mean(pred_img_iteration, dim=0, keepdim=True)                 # :423
    pred_img_var = torch.std(pred_img_iteration, dim=0, keepdim=True)              # :424
    noise_removal = pred_img_var > noise_threshold * pred_img                      # :425
    pred_img[noise_removal] = 0                                                    # :427
    out, std = mc_dropout_aggregate(pred_img_iteration.numpy(), noise_threshold)
    np.testing.assert_allclose(std, pred_img_var.numpy(), rtol=2e-5, atol=2e-6)   # identical passes: two-pass std is an ulp of the mean, Welford gives 0
    keep = np.abs(pred_img_var.numpy() - noise_threshold * torch.mean(pred_img_iteration, 0, keepdim=True).numpy()) > 1e-6
    np.testing.assert_allclose(out[keep], pred_img.numpy()[keep], rtol=1e-6, atol=1e-7)
    assert (out == 0).any() and (out != 0).any()
