"""Pins oracle/adamw.py against torch.optim.AdamW (the implementation the reference calls, main_lidar_upsampling.py:283)."""
import numpy as np
import torch

from oracle import adamw as A


def test_adamw_oracle_matches_torch():
    g = torch.Generator().manual_seed(3)
    shapes = [(96, 96), (96,), (288, 96)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    groups = [{"params": [params[0], params[2]], "weight_decay": 0.05, "lr": 5e-4}, {"params": [params[1]], "weight_decay": 0.0, "lr": 2.5e-4}]
    opt = torch.optim.AdamW(groups, lr=5e-4, betas=(0.9, 0.95), foreach=False)
    mine = [p.detach().numpy().copy() for p in params]
    m = [np.zeros_like(x) for x in mine]
    v = [np.zeros_like(x) for x in mine]
    hyp = {0: (5e-4, 0.05), 2: (5e-4, 0.05), 1: (2.5e-4, 0.0)}
    for step in range(1, 5):
        grads = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        want_norm = torch.norm(torch.stack([torch.norm(p.grad.detach(), 2.0) for p in params]), 2.0).item()     # misc.py:325-328
        assert abs(A.grad_norm([gr.numpy() for gr in grads]) - want_norm) <= 1e-5 * want_norm
        opt.step()
        for i in range(3):
            A.adamw_step(mine[i], grads[i].numpy(), m[i], v[i], step, hyp[i][0], 0.9, 0.95, 1e-8, hyp[i][1])
            np.testing.assert_allclose(mine[i], params[i].detach().numpy(), rtol=2e-6, atol=1e-8)
