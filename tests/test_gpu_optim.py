"""Fused AdamW / gradient norm over the flat buffers against torch.optim.AdamW -- the optimizer the reference
constructs (main_lidar_upsampling.py:283) -- on identical parameters and gradients, with layer-decay style groups."""
import numpy as np
import pytest
import torch

from oracle import adamw as A
from oracle.params import TULIP_BASE, make_inputs, make_params
from tests.test_gpu_model import build, load_params

pytestmark = pytest.mark.gpu


def groups_of(model, lr, wd):
    """two-by-two groups in the spirit of param_groups_layer_decay: no decay on 1-D tensors, a different lr for the decoder."""
    out = {}
    for n, p in model.named_parameters():
        key = (p.ndim == 1 or n.endswith(".bias"), n.startswith("layers_up"))
        out.setdefault(key, []).append(p)
    return [{"params": ps, "weight_decay": 0.0 if no_decay else wd, "lr": lr * (0.75 if dec else 1.0)} for (no_decay, dec), ps in out.items()]


def test_flat_adamw_matches_torch_adamw():
    from tulip_b200.optim import FlatAdamW
    cfg = TULIP_BASE
    pn = make_params(cfg, 51)
    lo, hi = make_inputs(cfg, 2, 52)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg).train()
    load_params(model, pn)
    model.cuda()
    ref_params = {n: torch.nn.Parameter(p.detach().cpu().clone()) for n, p in model.named_parameters()}
    names = {id(p): n for n, p in model.named_parameters()}
    g_mine = groups_of(model, 5e-4, 0.05)
    g_ref = [{"params": [ref_params[names[id(p)]] for p in g["params"]], "weight_decay": g["weight_decay"], "lr": g["lr"]} for g in g_mine]
    opt = FlatAdamW(model, g_mine, lr=5e-4, betas=(0.9, 0.95))
    ref = torch.optim.AdamW(g_ref, lr=5e-4, betas=(0.9, 0.95), foreach=False)
    for step in range(3):
        model.zero_grad()
        _, loss, _ = model(lo_t, hi_t)
        loss.backward()
        grads = {n: p.grad.detach().cpu().clone() for n, p in model.named_parameters()}
        want_norm = A.grad_norm([g.numpy() for g in grads.values()])
        assert abs(opt.grad_norm().item() - want_norm) <= 1e-5 * want_norm
        for n, p in ref_params.items():
            p.grad = grads[n]
        if step == 1:                                         # the scheduler of the reference rewrites lr every iteration
            for ga, gb in zip(opt.param_groups, ref.param_groups):
                ga["lr"] *= 0.5; gb["lr"] *= 0.5
        opt.step()
        ref.step()
        worst = 0.0
        for n, p in model.named_parameters():
            a, b = p.detach().cpu(), ref_params[n].detach()
            worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-12)))
        assert worst <= 5e-6, (step, worst)                   # fp32 element-wise update: 1e-6-level, contraction order only
    # parameters stayed views of the flat buffer and the next forward sees the update
    assert model._is_flat(torch.device("cuda", 0))
    _, loss2, _ = model(lo_t, hi_t)
    assert torch.isfinite(loss2) and loss2.item() != loss.item()


def test_flat_adamw_grad_scale_and_state_dict():
    """grad_scale multiplies the gradients inside the update exactly like unscaling the buffer first (GradScaler.unscale_)."""
    from tulip_b200.optim import FlatAdamW
    from tulip_b200.parallel import flat_grad_of
    cfg = TULIP_BASE
    lo, hi = make_inputs(cfg, 1, 53)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg).eval()
    load_params(model, make_params(cfg, 54))
    model.cuda()
    _, loss, _ = model(lo_t, hi_t)
    (loss * 1024.0).backward()                                  # GradScaler-style scaled loss
    start = model._flat.clone()
    opt1 = FlatAdamW(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01)
    opt1.step(grad_scale=1.0 / 1024.0)                          # unscaled on the fly inside the update
    after1 = model._flat.clone()
    model._flat.copy_(start)
    flat_grad_of(model).mul_(1.0 / 1024.0)                      # what GradScaler.unscale_ does
    opt2 = FlatAdamW(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01)
    opt2.step()
    assert torch.equal(model._flat, after1)
    assert not torch.equal(after1, start)
    sd = opt1.state_dict()
    opt3 = FlatAdamW(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.01)
    opt3.load_state_dict(sd)
    assert opt3.steps == 1 and torch.equal(opt3.exp_avg, opt1.exp_avg) and torch.equal(opt3.exp_avg_sq, opt2.exp_avg_sq)


def test_flat_adamw_state_dict_interchanges_with_torch_adamw():
    """a checkpoint written by torch.optim.AdamW (what misc.save_model stores, util/misc.py:338-344) must resume
    under FlatAdamW with its moments and step count, and the other way round."""
    from tulip_b200.optim import FlatAdamW
    cfg = TULIP_BASE
    lo, hi = make_inputs(cfg, 2, 52)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg).train()
    load_params(model, make_params(cfg, 51))
    model.cuda()
    ref = torch.optim.AdamW(groups_of(model, 5e-4, 0.05), lr=5e-4, betas=(0.9, 0.95))
    for _ in range(2):
        model.zero_grad()
        _, loss, _ = model(lo_t, hi_t)
        loss.backward()
        ref.step()
    sd = ref.state_dict()
    opt = FlatAdamW(model, groups_of(model, 5e-4, 0.05), lr=5e-4, betas=(0.9, 0.95))
    opt.load_state_dict(sd)
    assert opt.steps == 2
    k0 = next(iter(sd["state"]))
    p0 = ref.param_groups[0]["params"][0]
    o, n, shape = model._views[[id(p) for p in model._param_list].index(id(p0))]
    assert torch.equal(opt.exp_avg[o:o + n].view(shape), sd["state"][k0]["exp_avg"])
    # ... and back: torch.optim.AdamW resumes from FlatAdamW's state_dict and both take the same third step
    back = torch.optim.AdamW(groups_of(model, 5e-4, 0.05), lr=5e-4, betas=(0.9, 0.95))
    back.load_state_dict(opt.state_dict())
    assert int(float(back.state[p0]["step"])) == 2 and torch.equal(back.state[p0]["exp_avg_sq"], sd["state"][k0]["exp_avg_sq"])
    with pytest.raises(ValueError):
        bad = {"state": {0: {"foo": 1}}, "param_groups": sd["param_groups"]}
        FlatAdamW(model, groups_of(model, 5e-4, 0.05), lr=5e-4, betas=(0.9, 0.95)).load_state_dict(bad)
