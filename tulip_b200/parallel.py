"""Data-parallel plumbing: one process per GPU, frames sharded across ranks, ONE gradient all-reduce.

The reference wraps the model in torch DDP (main_lidar_upsampling.py:276-278), which reduces ~5 buckets of
fp32 gradients during backward.  Frames are independent on the TULIP path (no op mixes the batch dimension,
SURVEY.md 8e), and this implementation already owns every gradient in one flat fp32 buffer, so the whole
exchange step is a single NCCL all-reduce over NVLink followed by a 1/world scale.  (Under the unchanged
reference driver the module is still DDP-wrappable: its parameters are ordinary nn.Parameters.)
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, even split of a global batch; the reference's DistributedSampler + drop_last semantics
    (main_lidar_upsampling.py:172-178, 202-208) require n_items % world == 0."""
    if n_items % world:
        raise ValueError(f"global batch {n_items} is not divisible by world size {world}")
    per = n_items // world
    return range(rank * per, (rank + 1) * per)


def allreduce_flat_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean of a flat gradient buffer over the process group (DDP semantics: sum / world)."""
    if not dist.is_available() or not dist.is_initialized():
        return flat_grad
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grad
    if flat_grad.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)     # mean inside NCCL: no second pass over 108 MB
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.mul_(1.0 / world)
    return flat_grad


def flat_grad_of(model) -> torch.Tensor:
    """The flat fp32 buffer that every `.grad` of a tulip_b200 TULIP module currently aliases."""
    plist, views = model._param_list, model._views
    g0 = plist[0].grad
    if g0 is None:
        raise RuntimeError("no gradients: call backward() first")
    for buf in model._grad_bufs:
        if buf is not None and g0.data_ptr() == buf.data_ptr() + 4 * views[0][0]:
            last, (o, n, _s) = plist[-1].grad, views[-1]
            if last is not None and last.data_ptr() == buf.data_ptr() + 4 * o:
                return buf
    raise RuntimeError("gradients are not views of the flat buffer (were they re-assigned?)")


def allreduce_gradients(model, group=None) -> torch.Tensor:
    """The data-parallel exchange step: one all-reduce of all 27.1 M (tulip_base) gradients."""
    return allreduce_flat_(flat_grad_of(model), group)
