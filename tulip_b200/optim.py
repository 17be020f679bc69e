"""Fused optimizer step over the flat parameter / gradient buffers: the reference trains with
`torch.optim.AdamW(param_groups_layer_decay(model, wd), lr, betas=(0.9, 0.95))` and measures the gradient norm with
`get_grad_norm_` (main_lidar_upsampling.py:281-283, util/misc.py:294-329) -- a multi-tensor apply over 212 tensors plus a
212-term norm.  All parameters of the tulip_b200 module are views of one buffer, so both become one kernel launch."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, current_stream, load_library, ptr
from .parallel import flat_grad_of

MAX_GROUPS = 64


class _Hyper(C.Structure):                       # mirrors tulip_adamw_hyper (include/tulip_b200.h)
    _fields_ = [("lr", C.c_float * MAX_GROUPS), ("weight_decay", C.c_float * MAX_GROUPS), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("bias_correction1", C.c_float), ("bias_correction2_sqrt", C.c_float), ("grad_scale", C.c_float)]


_SEG_DTYPE = np.dtype([("offset", "<i8"), ("numel", "<i8"), ("group", "<i4"), ("pad", "<i4")])       # tulip_adamw_segment


class FlatAdamW(torch.optim.Optimizer):
    """Drop-in for `torch.optim.AdamW(params_or_groups, lr, betas, eps, weight_decay)` on a tulip_b200 TULIP module (same update,
    same param_groups interface, so lr schedulers that rewrite `group["lr"]` every iteration keep working).  All groups must share
    betas and eps (the reference passes them once)."""

    def __init__(self, model, params=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(list(model.parameters()) if params is None else params, defaults)
        if len(self.param_groups) > MAX_GROUPS:
            raise ValueError(f"FlatAdamW supports at most {MAX_GROUPS} parameter groups")
        self.model = model
        self.steps = 0
        self._flat_id = None

    def _prepare(self):
        model = self.model
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on CUDA only (tulip_b200 has no CPU path)")
        model._ensure_flat(dev)
        if self._flat_id == (model._flat.data_ptr(), id(model._views)):
            return
        index = {id(p): i for i, p in enumerate(model._param_list)}
        segs = []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if id(p) not in index:
                    raise ValueError("FlatAdamW: a parameter of the optimizer does not belong to the model")
                o, n, _ = model._views[index[id(p)]]
                segs.append((o, n, gi, 0))
        segs.sort()
        arr = np.array(segs, dtype=_SEG_DTYPE)
        self._segs = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        self._n_segs = len(segs)
        self._span = int(model._flat.numel())
        if getattr(self, "exp_avg", None) is None or self.exp_avg.shape != model._flat.shape or self.exp_avg.device != dev:
            self.exp_avg = torch.zeros_like(model._flat)
            self.exp_avg_sq = torch.zeros_like(model._flat)
        self._norm_scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        self._flat_id = (model._flat.data_ptr(), id(model._views))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One AdamW update of every parameter from the module's flat gradient buffer (`grad_scale` multiplies the gradients on the
        fly, e.g. 1 / GradScaler scale when they have not been unscaled)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._prepare()
        gbuf = flat_grad_of(self.model)
        self.steps += 1
        beta1, beta2 = self.param_groups[0]["betas"]
        hp = _Hyper()
        for gi, group in enumerate(self.param_groups):
            if tuple(group["betas"]) != (beta1, beta2) or group["eps"] != self.param_groups[0]["eps"]:
                raise ValueError("FlatAdamW: all groups must share betas and eps")
            hp.lr[gi], hp.weight_decay[gi] = group["lr"], group["weight_decay"]
        hp.beta1, hp.beta2, hp.eps = beta1, beta2, self.param_groups[0]["eps"]
        hp.bias_correction1 = 1.0 - beta1 ** self.steps
        hp.bias_correction2_sqrt = math.sqrt(1.0 - beta2 ** self.steps)
        hp.grad_scale = grad_scale
        check(load_library().tulip_adamw_step(ptr(self.model._flat), ptr(gbuf), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self._segs),
                                              self._n_segs, self._span, C.byref(hp), current_stream()), "tulip_adamw_step")
        self.model._param_epoch = getattr(self.model, "_param_epoch", 0) + 1     # the kernel wrote the parameters behind autograd's back
        return loss

    @torch.no_grad()
    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of all gradients (util/misc.py:317-329 with norm_type = 2) as a device scalar."""
        self._prepare()
        out = torch.empty(1, dtype=torch.float32, device=self._segs.device)
        check(load_library().tulip_grad_norm(ptr(flat_grad_of(self.model)), ptr(self._segs), self._n_segs, self._span, ptr(self._norm_scratch),
                                             ptr(out), current_stream()), "tulip_grad_norm")
        return out[0]

    def _packed_params(self):
        """(packed index, parameter) pairs in the order torch.optim.Optimizer.state_dict numbers them."""
        out, idx = [], 0
        for group in self.param_groups:
            for p in group["params"]:
                out.append((idx, p))
                idx += 1
        return out

    def state_dict(self):
        """The standard torch.optim.AdamW layout: state[i] = {step, exp_avg, exp_avg_sq} per parameter (the moments are views of
        the flat buffers), so a checkpoint written by `misc.save_model` (util/misc.py:332-349) resumes under either optimizer."""
        sd = super().state_dict()
        if self.steps > 0 and getattr(self, "exp_avg", None) is not None:
            index = {id(p): i for i, p in enumerate(self.model._param_list)}
            state = {}
            for k, p in self._packed_params():
                o, n, shape = self.model._views[index[id(p)]]
                state[k] = {"step": torch.tensor(float(self.steps)), "exp_avg": self.exp_avg[o:o + n].view(shape),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(shape)}
            sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        """Accepts the standard per-parameter AdamW state (a reference / torch.optim.AdamW checkpoint) and the round-1 private
        `flat` entry.  A state that holds neither for a non-empty optimizer raises instead of silently restarting the moments."""
        flat = state_dict.get("flat")
        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        if flat is not None:
            self.steps = int(flat["steps"])
            if flat["exp_avg"] is not None:
                self._prepare()
                self.exp_avg.copy_(flat["exp_avg"]); self.exp_avg_sq.copy_(flat["exp_avg_sq"])
            return
        if not state:
            self.steps = 0
            if getattr(self, "exp_avg", None) is not None:
                self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            return
        self._prepare()
        index = {id(p): i for i, p in enumerate(self.model._param_list)}
        steps = set()
        for k, p in self._packed_params():
            st = state.get(k, state.get(str(k)))
            if st is None or "exp_avg" not in st or "exp_avg_sq" not in st or "step" not in st:
                raise ValueError(f"FlatAdamW.load_state_dict: no AdamW state (step / exp_avg / exp_avg_sq) for parameter {k}")
            o, n, shape = self.model._views[index[id(p)]]
            self.exp_avg[o:o + n].view(shape).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + n].view(shape).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise ValueError(f"FlatAdamW.load_state_dict: parameters disagree on the step count ({sorted(steps)})")
        self.steps = steps.pop()
