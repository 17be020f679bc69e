"""world_size-2 gloo test of the data-parallel host logic (shard + single flat all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tulip_b200.parallel import allreduce_flat_, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    allreduce_flat_(g)
    want = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(g, want)
    # frame sharding: the union of the ranks' ranges is the global batch, disjoint and ordered
    r = shard_range(32, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, list(r))
    flat = [i for part in gathered for i in part]
    ok = ok and flat == list(range(32))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_allreduce_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_rejects_ragged():
    with pytest.raises(ValueError):
        shard_range(33, 0, 2)
    assert list(shard_range(8, 1, 4)) == [2, 3]
    t = torch.ones(4)
    assert allreduce_flat_(t) is t          # no process group: identity


def test_phase_slices_partition_the_flat_buffer():
    """tulip_b200.parallel.phase_slices: the three backward phases' gradient ranges are disjoint, cover every parameter and follow
    the order in which tulip_net_backward_phases completes them (head + decoder, top encoder stage, rest + PatchEmbed)."""
    import numpy as np
    from oracle.params import TULIP_BASE, TULIP_LARGE, param_shapes
    from tulip_b200.parallel import phase_slices
    for cfg in (TULIP_BASE, TULIP_LARGE):
        shapes = param_shapes(cfg)
        names = [k for k in shapes if not k.endswith("relative_position_index")]
        views, off = [], 0
        for n in names:
            k = int(np.prod(shapes[n]))
            views.append((off, k, shapes[n]))
            off += (k + 63) // 64 * 64
        sl = phase_slices(names, views, cfg.num_layers)
        covered = np.zeros(off, dtype=np.int8)
        for ph in range(3):
            for lo, hi in sl[ph]:
                assert 0 <= lo < hi <= off
                covered[lo:hi] += 1
        assert covered.max() == 1 and covered.min() == 1
        top = f"layers.{cfg.num_layers - 1}."
        for n, (o, k, _) in zip(names, views):
            ph = [i for i in range(3) if any(lo <= o and o + k <= hi for lo, hi in sl[i])]
            want = 2 if n.startswith("patch_embed.") else (1 if n.startswith(top) else (2 if n.startswith("layers.") else 0))
            assert ph == [want], n


def test_overlap_switch_records_group_and_sm_reserve():
    """overlap_gradient_allreduce only flips host-side switches the autograd node reads (tulip_b200/model/tulip.py backward):
    the process group (True = default group), the SM reserve handed to tulip_set_sm_budget, and no pending work."""
    import types
    from tulip_b200.parallel import overlap_gradient_allreduce
    m = types.SimpleNamespace()
    assert overlap_gradient_allreduce(m) is m and m._grad_sync is True and m._grad_sync_reserve_sms == 0
    overlap_gradient_allreduce(m, group="g", reserve_sms=8)
    assert m._grad_sync == "g" and m._grad_sync_reserve_sms == 8 and m._grad_sync_pending is None
    overlap_gradient_allreduce(m, enabled=False, reserve_sms=8)
    assert m._grad_sync is None and m._grad_sync_reserve_sms == 0
