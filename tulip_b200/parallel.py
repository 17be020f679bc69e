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
    """The data-parallel exchange step: one all-reduce of all 27.1 M (tulip_base) gradients.  A model with
    `overlap_gradient_allreduce` enabled has already exchanged them inside backward(): then this only waits for that."""
    pending = getattr(model, "_grad_sync_pending", None)
    if pending is not None:
        for w in pending:
            w.wait()                                   # stream-level wait on the NCCL stream, not a host sync
        model._grad_sync_pending = None
        return flat_grad_of(model)
    return allreduce_flat_(flat_grad_of(model), group)


def phase_slices(names, views, num_layers: int):
    """Element ranges of the flat gradient buffer that are complete after backward phase 0, 1 and 2 (tulip_net_backward_phases):
    -> ([(lo, hi), ...] per phase).  `names` / `views` are the module's parameter names and (offset, numel, shape) triples in
    flat-buffer order (state_dict order: layers.*, layers_up.*, first_patch_expanding, skip_connection_layers, norm_up,
    patch_embed, decoder_pred, ps_head)."""
    top = f"layers.{num_layers - 1}."

    def phase_of(n):
        if n.startswith("patch_embed."):
            return 2
        if n.startswith("layers."):
            return 1 if n.startswith(top) else 2
        return 0
    end = max(o + k for o, k, _ in views)
    bounds = [o for o, _, _ in views] + [end]
    out = ([], [], [])
    for i, n in enumerate(names):
        ph, lo, hi = phase_of(n), bounds[i], bounds[i + 1]
        if out[ph] and out[ph][-1][1] == lo:
            out[ph][-1] = (out[ph][-1][0], hi)         # contiguous with the previous range of the same phase (padding included)
        else:
            out[ph].append((lo, hi))
    return out


def overlap_gradient_allreduce(model, group=None, enabled: bool = True, reserve_sms: int = 0):
    """Exchange the gradients INSIDE backward(): the pass runs as three phases (head + decoder, top encoder stage, the rest) and
    the all-reduce of each finished slice of the flat buffer is launched on NCCL's stream under the remaining phases -- what the
    reference gets from DDP's bucketed reduce (main_lidar_upsampling.py:277).  Same arithmetic as one flat all-reduce (mean over
    ranks of every element), 4 collectives instead of 1.  Call `allreduce_gradients(model)` after backward() as before: it then
    only joins the NCCL stream.  Do not combine with DistributedDataParallel or with gradient accumulation across backward() calls.
    `reserve_sms` > 0: once the first reduce is in flight the persistent kernels of the remaining phases are launched for
    (SMs - reserve_sms) SMs (tulip_set_sm_budget), so their CTAs do not queue behind NCCL's; pair it with NCCL_MAX_CTAS."""
    model._grad_sync = (group if group is not None else True) if enabled else None
    model._grad_sync_reserve_sms = int(reserve_sms) if enabled else 0
    model._grad_sync_pending = None
    return model


def launch_slice_allreduce(gbuf: torch.Tensor, slices, group) -> list:
    """async mean all-reduce of element ranges of the flat gradient buffer; returns the work handles"""
    works = []
    nccl = gbuf.is_cuda and dist.get_backend(group) == "nccl"
    for lo, hi in slices:
        t = gbuf[lo:hi]
        if nccl:
            works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True))
        else:
            w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True)
            w.wait()
            t.mul_(1.0 / dist.get_world_size(group))
    return works
