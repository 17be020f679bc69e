"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the optimizer update the reference trains with: torch.optim.AdamW (decoupled
weight decay, no amsgrad; main_lidar_upsampling.py:283) in the order of torch's _single_tensor_adamw, and get_grad_norm_
(util/misc.py:317-329).  Pinned by tests/test_oracle_adamw.py against torch.optim.AdamW itself on CPU."""
import math

import numpy as np


def adamw_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay):
    """in-place on float32 arrays; `step` is the 1-based step count."""
    f = np.float32
    p *= f(1 - lr * weight_decay)
    m += (g - m) * f(1 - beta1)
    v *= f(beta2)
    v += f(1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2_sqrt = math.sqrt(1 - beta2 ** step)
    denom = np.sqrt(v) / f(bc2_sqrt) + f(eps)
    p -= f(lr / bc1) * (m / denom)


def grad_norm(grads):
    return math.sqrt(sum(float((np.asarray(g, np.float64) ** 2).sum()) for g in grads))
