#!/usr/bin/env python
"""Benchmark of the TULIP Swin forward/backward hot path (BASELINE.json metric: range-image frames/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- tulip_base, KITTI 16x1024 -> 64x1024, batch 32 per GPU,
one training step = forward + L1 loss + backward (+ ONE flat-gradient NCCL all-reduce when N > 1).  The metric is
"fwd+bwd", so the optimizer update is outside the timed step (its cost is reported separately as `adamw_ms`).
One rank per GPU, weights replicated, frames sharded (weak scaling: 32 frames per GPU).

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same step driven through the
public nn.Module API with inputs coming from pinned host memory (H2D copies and the loss read-back inside the timed
region).  `roofline` describes the dominant kernel function (per-launch CUDA events recorded by the C executor's
built-in profiler over extra steps after the timed region); `cpu_baseline` is the oracle's CPU port of the reference
path timed on this box's host cores on a bounded sample.  `--impl reference` times that CPU port alone.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "range-image frames/sec (fwd+bwd)"
UNIT = "frames/s"
COMMON_KW = dict(patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False, pixel_shuffle=True, circular_padding=True,
                 log_transform=True, patch_unmerging=True)
# BASELINE.json configs[1], [2] and [4]; FLOPs per frame (fwd+bwd, 2 x MAC, matmul/conv) from SURVEY.md 8d.  `wide` is the
# cfg5 surrogate of SURVEY 8d option (i): the reference cannot build "embed_dim=192, depths=[2,2,18,2], 8x" (its upscale
# formula gives r = 4), so the same trunk runs at 16x1024 -> 64x1024 through the TULIP constructor.
CONFIGS = {
    "kitti32": dict(label="KITTI 16x1024->64x1024 tulip_base", img=(16, 1024), tgt=(64, 1024), batch=32, arch="base",
                    flops=46_349_156_352, wmsa_fwd_flops=4_410_310_656),
    "durlar16": dict(label="DurLAR 32x2048->128x2048 tulip_base", img=(32, 2048), tgt=(128, 2048), batch=16, arch="base",
                     flops=185_396_625_408, wmsa_fwd_flops=4 * 4_410_310_656),
    "large8": dict(label="cfg5 surrogate: TULIP(embed_dim=192, depths=(2,2,18,2)) KITTI 16x1024->64x1024", img=(16, 1024),
                   tgt=(64, 1024), batch=8, arch="wide", flops=533_301_559_296, wmsa_fwd_flops=None),
}
CFG = CONFIGS["kitti32"]
IMG, TGT = CFG["img"], CFG["tgt"]


def select_config(name):
    global CFG, IMG, TGT
    CFG = CONFIGS[name]
    IMG, TGT = CFG["img"], CFG["tgt"]


def build_model(mod):
    """`mod` is a module with the reference's model API (tulip_b200.model.tulip or the reference's own model.tulip)."""
    kw = dict(img_size=IMG, target_img_size=TGT, **COMMON_KW)
    if CFG["arch"] == "base":
        return mod.tulip_base(**kw)
    from functools import partial
    return mod.TULIP(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48), mlp_ratio=4, qkv_bias=True, drop_rate=0,
                     attn_drop_rate=0, drop_path_rate=0.1, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def synth_inputs(batch, seed):
    """lo = log1p(U[0,1)) (B,1,16,1024), hi likewise (B,1,64,1024), 15 % invalid pixels zeroed."""
    g = torch.Generator().manual_seed(seed)
    hi = torch.log1p(torch.rand((batch, 1, *TGT), generator=g))
    hi = torch.where(torch.rand(hi.shape, generator=g) < 0.15, torch.zeros(()), hi)
    lo = hi[:, :, ::TGT[0] // IMG[0], :].contiguous()
    return lo, hi


# ----------------------------------------------------------------------------------------------- CPU reference arm
def reference_module():
    """The UNMODIFIED reference (baseline/_ref, staged by __graft_entry__.build()) or None where it has not been staged."""
    try:
        from compat.env import import_reference_model, reference_available
        if reference_available():
            return import_reference_model()
    except Exception as e:                                  # pragma: no cover
        print(f"bench.py: reference not importable ({e}); falling back to the oracle port", file=sys.stderr)
    return None


def cpu_frames_per_s(batch, reps, warm, threads):
    """Reference path on the host cores: the reference's own torch-eager model (kind "reference") when baseline/_ref is there,
    else the oracle's functional fp32 port of it (kind "port").  fp32, train mode, forward + L1 + backward."""
    torch.set_num_threads(threads)
    lo, hi = synth_inputs(batch, 1)
    T = reference_module()
    if T is not None:
        torch.manual_seed(0)
        model = build_model(T).train()

        def one():
            model.zero_grad(set_to_none=True)
            _, loss, _ = model(lo, hi, eval=False)
            loss.backward()
        kind = "reference"
    else:
        from oracle import tulip_oracle as O
        from oracle.params import Cfg, make_params
        ocfg = Cfg(img_size=IMG, target_img_size=TGT) if CFG["arch"] == "base" else \
            Cfg(img_size=IMG, target_img_size=TGT, embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48))
        p = O.to_torch(make_params(ocfg, 0), requires_grad=True)

        def one():
            for q in p.values():
                q.grad = None
            _, loss, _ = O.forward(p, ocfg, lo, hi)
            loss.backward()
        kind = "port"
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return batch / (sum(times) / len(times)), times, kind


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # size the per-step sample so the whole run stays within a few minutes
    fps1, _, _ = cpu_frames_per_s(1, 1, 1, threads)
    per_frame = 1.0 / fps1
    budget = 150.0
    frames = int(max(1, min(CFG["batch"], budget / ((args.steps + args.warmup) * per_frame))))
    fps, times, kind = cpu_frames_per_s(frames, args.steps, args.warmup, threads)
    ms = 1e3 * sum(times) / len(times)
    what = "the reference's own torch-eager model (baseline/_ref)" if kind == "reference" else "CPU torch-eager port (oracle)"
    out = {
        "impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CFG['label']} fwd+L1+bwd, {what} on the host cores, {frames} frames per step",
                   "frames_per_step": frames, "config_name": args.config},
        "cpu_baseline": {"value": round(fps, 3), "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steps x {frames} frames, fp32, torch {torch.__version__} eager on {threads} threads"},
        "e2e": {"value": round(fps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def gpu_eager_baseline(B, dev, steps=10, warm=3):
    """The north star's >= 10x denominator: the UNMODIFIED reference model run by torch eager on this GPU, same workload and
    batch, forward + L1 + backward in train mode, under torch.autocast (bf16, and the engine's native fp16 of
    engine_upsampling.py:77).  Device-resident inputs, CUDA events.  None where baseline/_ref has not been staged."""
    T = reference_module()
    if T is None:
        return None
    torch.manual_seed(0)
    model = build_model(T).to(dev).train()
    lo, hi = (t.to(dev) for t in synth_inputs(B, 1))
    out = {"what": "reference tulip/model/tulip.py (baseline/_ref, unmodified), torch eager on the same GPU, fwd+L1+bwd, train mode",
           "batch": B, "torch": torch.__version__, "cudnn_benchmark": True}
    torch.backends.cudnn.benchmark = True                  # main_lidar_upsampling.py:159
    for name, dt in (("autocast_bf16", torch.bfloat16), ("autocast_fp16", torch.float16)):
        scaler = torch.amp.GradScaler("cuda", enabled=(dt == torch.float16))

        def one():
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=dt):
                _, loss, _ = model(lo, hi, eval=False)
            scaler.scale(loss).backward()
        for _ in range(warm):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": round(ms, 3), "value": round(B / (ms * 1e-3), 2), "unit": UNIT}
    del model
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region.  NVML polled every 5 ms from a thread (the timed region
    of a default run is ~0.1 s, too short for `nvidia-smi -lms`); falls back to an nvidia-smi loop if pynvml is unusable."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.p = self.f = self.thread = None
        self.samples = []
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            sm = [s[0] for s in self.samples]
            mask = 0
            for s in self.samples:
                mask |= s[2]
            return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_mhz, "power_w_max": float(max(s[1] for s in self.samples)),
                    "samples": len(sm), "reasons": sorted(k for k, b in self.BITS.items() if mask & b), "source": "nvml, 5 ms poll"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def read_profile(model):
    from tulip_b200._lib import load_library
    lib = load_library()
    out = []
    name = C.create_string_buffer(64)
    ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    for t in range(lib.tulip_net_profile_num_tags()):
        rc = lib.tulip_net_profile_read(model._net, t, name, 64, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n))
        if rc == 0 and n.value > 0:
            out.append({"kernel": name.value.decode(), "ms": ms.value, "flops": fl.value, "bytes": by.value, "launches": n.value})
    return out


def read_records(model):
    """per-launch records in launch order: (kernel, stage, part, backward, ms, flops, bytes)"""
    from tulip_b200._lib import load_library
    lib = load_library()
    names = {}
    nm = C.create_string_buffer(64)
    ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    for t in range(lib.tulip_net_profile_num_tags()):
        lib.tulip_net_profile_read(model._net, t, nm, 64, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n))
        names[t] = nm.value.decode()
    tag, st, part, bwd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    out = []
    cnt = lib.tulip_net_profile_record(model._net, 0, C.byref(tag), C.byref(ms), C.byref(fl), C.byref(by))
    for i in range(max(cnt, 0)):
        lib.tulip_net_profile_record(model._net, i, C.byref(tag), C.byref(ms), C.byref(fl), C.byref(by))
        lib.tulip_net_profile_where(model._net, i, C.byref(st), C.byref(part), C.byref(bwd))
        out.append((names.get(tag.value, "?"), st.value, part.value, bwd.value, ms.value, fl.value, by.value))
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from tulip_b200._lib import load_library
    from tulip_b200.parallel import allreduce_gradients, overlap_gradient_allreduce

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (tulip_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = load_library()
    import tulip_b200.model.tulip as tb_model
    B = args.batch
    torch.manual_seed(0)                                   # identical replicas on every rank (reference: DDP broadcast)
    model = build_model(tb_model).to(dev).train()          # train mode: DropPath masks are drawn every step, as in the reference
    torch.manual_seed(0 + rank)                            # per-rank RNG stream for the DropPath masks (reference: seed + rank, main:155)
    if world > 1 and args.allreduce == "overlap":
        # slices of the flat buffer are reduced under the rest of backward, which leaves NCCL's SMs alone (--reserve-sms)
        overlap_gradient_allreduce(model, reserve_sms=args.reserve_sms)
    lo_h, hi_h = synth_inputs(B, 1 + rank)
    lo_pin, hi_pin = lo_h.pin_memory(), hi_h.pin_memory()
    lo_d, hi_d = lo_pin.to(dev), hi_pin.to(dev)

    def local_step(lo, hi):
        _, loss, _ = model(lo, hi)
        loss.backward()
        return loss

    def step(lo, hi):
        loss = local_step(lo, hi)
        if world > 1:
            allreduce_gradients(model)       # the one exchange step of the path: a single flat NCCL all-reduce
        # gradients are dropped where the reference loop calls optimizer.zero_grad(): at the end of the iteration
        # (engine_upsampling.py:97-98), i.e. while the GPU still runs this step, not between the loss read and the next forward
        model.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step(lo_d, hi_d)
    launches0 = model.kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms = timed(lambda: step(lo_d, hi_d), args.steps)
    clocks = sampler.stop() if sampler else None
    launches = model.kernel_launches() - launches0
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # end to end: host-resident inputs in pinned memory -> H2D -> step -> loss read back, every step.  The input pipeline is
    # the usual double-buffered prefetcher: while step i computes, a copy stream uploads the batch of step i+1 from pinned
    # memory into the other device buffer; the loss of step i is read back (a host sync) before step i+1 is issued.
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [(torch.empty_like(lo_d), torch.empty_like(hi_d)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    slot = [0]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])            # the step that read this buffer has finished with it
            dbuf[i][0].copy_(lo_pin, non_blocking=True)
            dbuf[i][1].copy_(hi_pin, non_blocking=True)
            ready[i].record(copy_stream)

    def e2e_step():
        i = slot[0]
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[i])
        loss = step(dbuf[i][0], dbuf[i][1])
        consumed[i].record(cur)
        prefetch(i ^ 1)
        slot[0] = i ^ 1
        # the module's own loss read-back: (total, pixel) copied to pinned host memory behind the forward on a side stream, so
        # the host sync waits for this step's forward, as `loss.item()` does in the reference loop where it comes BEFORE
        # backward (engine_upsampling.py:84-93) -- not for the backward pass this loop has already queued
        return model.loss_item() if not args.e2e_item else loss.item()
    prefetch(0)
    for _ in range(2):
        e2e_step()
    sampler_e = ClockSampler(local_rank) if rank == 0 else None
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    clocks_e = sampler_e.stop() if sampler_e else None
    e2e_value = B * world / (e2e_ms * 1e-3)

    # (taken after the end-to-end leg, so that leg runs in the same power / thermal state as `value`, not behind 2.5 s at ~1 kW)
    # sustained reading: the same step for >= 2.5 s back to back with its own clock record (the driver's --steps 20 region is
    # ~0.1 s at boost clock; MEASURED_PEAKS' sustained figures were taken after seconds under load)
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2500.0 / ms_per_step) + 1)
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        sus_ms = timed(lambda: step(lo_d, hi_d), n_sus) / n_sus
        clocks2 = sampler2.stop() if sampler2 else None
        sustained = {"steps": n_sus, "seconds": round(sus_ms * n_sus * 1e-3, 3), "ms_per_step": round(sus_ms, 4),
                     "value": round(B * world / (sus_ms * 1e-3), 2), "unit": UNIT, "clocks": clocks2}

    if world > 1:
        overlap_gradient_allreduce(model, enabled=False)   # the sections below run on rank 0 alone: no collectives from here on
    if rank != 0:
        return
    # evaluation path (not part of the metric): forward-only + fused post-processing, as evaluate() runs it (B = 1)
    # and at B = 8, device-resident inputs, CUDA events
    from tulip_b200.inference import upsample
    DATASET = "durlar" if args.config == "durlar16" else "kitti"
    eval_path = {}
    model.eval()
    for eb in (1, 8):
        lo_e, hi_e = lo_d[:eb].contiguous(), hi_d[:eb].contiguous()
        for _ in range(5):
            upsample(model, lo_e, hi_e, DATASET)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            upsample(model, lo_e, hi_e, DATASET)
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1) / 30
        eval_path[f"b{eb}"] = {"ms_per_call": round(ms_e, 4), "frames_per_s": round(eb / (ms_e * 1e-3), 1)}
    model.train()
    # metric block of evaluate() for one frame: projection x2, Chamfer distance, voxel IoU / precision / recall
    from tulip_b200 import metrics as tb_metrics
    with torch.no_grad():
        img_p, _ = upsample(model, lo_d[:1].contiguous(), hi_d[:1].contiguous(), DATASET)
        img_g = torch.expm1(hi_d[:1]).contiguous()
        for _ in range(2):
            tb_metrics.evaluate_frame(img_p, img_g, DATASET, 0.1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            tb_metrics.evaluate_frame(img_p, img_g, DATASET, 0.1)
        eval_path["frame_metrics_ms"] = round((time.perf_counter() - t0) * 100.0, 4)
    # optimizer cost, outside the metric (torch fused AdamW over the 212 parameter views)
    # (rank 0 only from here on: no collectives)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), fused=True)
    local_step(lo_d, hi_d)
    for _ in range(2):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    adamw_ms = e0.elapsed_time(e1) / 5
    # the same update as ONE launch over the flat buffers (tulip_b200.optim.FlatAdamW)
    from tulip_b200.optim import FlatAdamW
    fopt = FlatAdamW(model, lr=1e-4, betas=(0.9, 0.95))
    for _ in range(2):
        fopt.step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fopt.step()
    e1.record()
    torch.cuda.synchronize()
    flat_adamw_ms = e0.elapsed_time(e1) / 5

    # per-kernel profile (extra steps, CUDA events around every launch on the launching stream)
    prof_steps = 3
    lib.tulip_net_profile(model._net, 1)
    for _ in range(prof_steps):
        local_step(lo_d, hi_d)
    torch.cuda.synchronize()
    prof = read_profile(model)
    records = read_records(model)
    lib.tulip_net_profile(model._net, 0)
    pk = peaks()
    tot = sum(k["ms"] for k in prof) or 1.0
    prof.sort(key=lambda k: -k["ms"])
    top = prof[0]
    per_launch_ms = top["ms"] / top["launches"]
    ridge = pk["tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    ai = top["flops"] / top["bytes"] if top["bytes"] else 0.0
    tflops = top["flops"] / (top["ms"] * 1e-3) / 1e12
    gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    if top["flops"] > 0 and ai >= ridge:
        roof = {"kernel": top["kernel"], "bound": "tensor", "achieved": round(tflops, 2), "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(tflops / pk["tflops_sustained"], 4)}
    else:
        # algorithmic intensity below the ridge (%.0f FLOP/B): the kernel is bounded by HBM bytes, not by the tensor pipe
        roof = {"kernel": top["kernel"], "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(gbs / pk["hbm_gbs"], 4)}
    roof.update({"arithmetic_intensity_flop_per_byte": round(ai, 1), "ridge_flop_per_byte": round(ridge, 1),
                 "tflops": round(tflops, 2), "tensor_frac": round(tflops / pk["tflops_sustained"], 4)})
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):                         # dram bytes per launch from a committed ncu capture (scripts/capture_traffic.sh)
        traffic = json.load(open(tpath)).get(top["kernel"], {}).get("dram_bytes_per_launch")
    roof.update({"traffic": traffic, "peak_source": pk["source"] + ", sustained figure for the tensor peak",
                 "avg_launch_ms": round(per_launch_ms, 4), "launches_per_step": top["launches"] // prof_steps,
                 "algorithmic_bytes_per_launch": round(top["bytes"] / top["launches"]),
                 "share_of_step": round(top["ms"] / tot, 4),
                 "kernels": [{"kernel": k["kernel"], "share": round(k["ms"] / tot, 4), "ms_per_step": round(k["ms"] / prof_steps, 4),
                              "launches_per_step": k["launches"] // prof_steps,
                              "tflops": round(k["flops"] / (k["ms"] * 1e-3) / 1e12, 2) if k["flops"] else None,
                              "gbs": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1)} for k in prof]})
    model_tflops = CFG["flops"] * B / (ms_per_step * 1e-3) / 1e12

    # per-stage entries of the dominant kernel function and of the whole step (one aggregate hides the stage split: stage 0 is
    # HBM-bound, the deep stages latency / operand-traffic bound)
    def agg(rows):
        ms_ = sum(r[4] for r in rows) / prof_steps
        fl_ = sum(r[5] for r in rows) / prof_steps
        by_ = sum(r[6] for r in rows) / prof_steps
        d = {"ms_per_step": round(ms_, 4), "launches_per_step": len(rows) // prof_steps}
        if ms_ > 0:
            d.update({"tflops": round(fl_ / (ms_ * 1e-3) / 1e12, 2), "gbs": round(by_ / (ms_ * 1e-3) / 1e9, 1),
                      "hbm_frac": round(by_ / (ms_ * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                      "tensor_frac": round(fl_ / (ms_ * 1e-3) / 1e12 / pk["tflops_sustained"], 4)})
        return d
    n_stages = 1 + max((r[1] for r in records), default=0)
    roof["per_stage"] = {f"stage{s_}": agg([r for r in records if r[0] == top["kernel"] and r[1] == s_]) for s_ in range(n_stages)}
    part_names = {0: "glue (embed / merge / unmerge / skip)", 1: "attention half-blocks", 2: "MLP half-blocks", 3: "head + loss"}
    step_split = {f"stage{s_}": {"fwd": agg([r for r in records if r[1] == s_ and r[3] == 0 and r[2] in (0, 1, 2)]),
                                 "bwd": agg([r for r in records if r[1] == s_ and r[3] == 1 and r[2] in (0, 1, 2)])}
                  for s_ in range(n_stages)}
    step_split["head"] = {"fwd": agg([r for r in records if r[2] == 3 and r[3] == 0]), "bwd": agg([r for r in records if r[2] == 3 and r[3] == 1])}
    step_split["parts"] = {part_names[p_]: agg([r for r in records if r[2] == p_]) for p_ in part_names}

    # W-MSA (the second half of BASELINE.json's metric): all attention half-block launches of the forward pass against SURVEY 8d's
    # W-MSA-kernel FLOPs per frame; the fused kernel's own launches; tensor-pipe % from the committed ncu capture
    wmsa = None
    att_fwd = [r for r in records if r[2] == 1 and r[3] == 0]
    if att_fwd and CFG["wmsa_fwd_flops"]:
        ms_att = sum(r[4] for r in att_fwd) / prof_steps
        tf_att = CFG["wmsa_fwd_flops"] * B / (ms_att * 1e-3) / 1e12
        wmsa = {"forward_ms_per_step": round(ms_att, 4), "flops_per_frame": CFG["wmsa_fwd_flops"], "tflops": round(tf_att, 2),
                "frac": round(tf_att / pk["tflops_sustained"], 4), "peak": pk["tflops_sustained"],
                "launches_per_step": len(att_fwd) // prof_steps,
                "note": "event-timed one launch at a time (no overlap between launches); training forward"}
    ncu_path = os.path.join(ROOT, "profiles", "wmsa_ncu.json")
    if os.path.exists(ncu_path):
        wmsa = dict(wmsa or {}, fused_kernel_ncu=json.load(open(ncu_path)))

    cpu = None
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps, times, kind = cpu_frames_per_s(4 if CFG["arch"] == "base" else 1, 3, 1, threads)
        cpu = {"value": round(fps, 3), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"3 steps x {4 if CFG['arch'] == 'base' else 1} frames of the same workload (fp32 torch eager, "
                         + ("the reference's own model from baseline/_ref" if kind == "reference" else "oracle port of the reference path")
                         + f"), {threads} threads"}
    if world == 1 and not args.no_eager_baseline:
        del model, opt, fopt                               # free the arena before the eager reference allocates its activations
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(B, dev)
        if eager is not None:
            for k in ("autocast_bf16", "autocast_fp16"):
                eager[k]["ours_over_eager"] = round(value / eager[k]["value"], 2)

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{CFG['label']}, batch {B}/GPU, train step = fwd + L1 + bwd"
                               + ((" + one flat NCCL grad all-reduce" if args.allreduce == "flat" else
                                   " + NCCL grad all-reduce of the flat buffer in 4 slices launched under backward") if world > 1 else ""),
                   "config_name": args.config, "global_batch": B * world, "parallelism": f"dp{world}", "optimizer": "outside the metric (fwd+bwd); see adamw_ms",
                   "l2": "activation working set of a step is GBs >> 126 MB L2 (no flush needed)", "train_mode": True},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(e2e_ms, 4),
                "h2d_bytes_per_step": int(lo_pin.numel() * 4 + hi_pin.numel() * 4),
                "d2h_bytes_per_step": 4 if args.e2e_item else 8, "clocks": clocks_e,
                "pipeline": "pinned host batch -> copy stream (double-buffered, uploads batch i+1 during step i) -> model(lo, hi); "
                            "loss.backward(); " + ("loss.item()" if args.e2e_item else "model.loss_item() (losses copied to pinned "
                            "host memory behind the forward; the host waits for that copy)") + " every step"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "step_split": step_split,
        "wmsa": wmsa,
        "sustained": sustained,
        "gpu_eager_baseline": eager,
        "eval_path": eval_path,
        "model_tflops": round(model_tflops, 2),
        "model_frac_of_tensor_roofline": round(model_tflops / pk["tflops_sustained"], 4),
        "adamw_ms": round(adamw_ms, 4), "flat_adamw_ms": round(flat_adamw_ms, 4),
        "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="kitti32", choices=sorted(CONFIGS),
                    help="kitti32 = BASELINE configs[1] (default, the metric's config); durlar16 = configs[2]; large8 = configs[4] surrogate")
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU (default: the config's batch)")
    ap.add_argument("--allreduce", default="flat", choices=["flat", "overlap"],
                    help="N > 1: one flat all-reduce after backward, or the same bytes in 4 slices launched under the backward phases")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="with --allreduce overlap: SMs left to NCCL (NCCL_MAX_CTAS) while the backward pass runs beside it")
    ap.add_argument("--e2e-item", action="store_true",
                    help="end-to-end loop reads the loss with loss.item() on the compute stream (waits for the queued backward) "
                         "instead of model.loss_item()")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the reference's GPU torch-eager arm (gpu_eager_baseline)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2.5 s sustained reading")
    args = ap.parse_args()
    select_config(args.config)
    if args.batch is None:
        args.batch = CFG["batch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.allreduce == "overlap":
            # NCCL's kernels must win SM slots against the persistent 148-CTA kernels of the backward pass they run under
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
            if args.reserve_sms > 0:
                os.environ.setdefault("NCCL_MAX_CTAS", str(args.reserve_sms))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
