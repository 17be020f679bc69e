"""Whole-model parity: the drop-in nn.Module (C++/CUDA executor) against the fp32 oracle and the fixtures
generated from the unmodified reference.

Tolerances.  Activations are bf16 between kernels with fp32 accumulation, so the end-to-end deviation from the
fp32 oracle is bounded by the reference's OWN autocast-bf16 deviation on the same kind of input
(9.1e-3 rel-L2 on pred with bf16-rounded weights; per-parameter gradient rel-L2 median 9.6e-3,
worst 0.23 on LayerNorm biases because the L1 loss back-propagates sign()).  The asserts below are:
pred rel-L2 <= 1e-2 (measured 3-4.5e-3), losses within 1e-3 relative, per-parameter gradient rel-L2 median <= 2e-2 and
every gradient norm within 15 % (LayerNorm / bias vectors 35 %)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import tulip_oracle as O
from oracle.params import TULIP_BASE, TULIP_LARGE, Cfg, make_inputs, make_params
from tests.util import rel_l2

pytestmark = pytest.mark.gpu

KW = dict(patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False, pixel_shuffle=True, circular_padding=True,
          log_transform=True, patch_unmerging=True)


def build(cfg, large=False):
    from functools import partial
    from tulip_b200.model.tulip import TULIP, tulip_base, tulip_large
    KW = {**globals()["KW"], "patch_unmerging": cfg.patch_unmerging, "pixel_shuffle": cfg.pixel_shuffle}
    if cfg.embed_dim != 96 or (tuple(cfg.depths) not in ((2, 2, 2, 2), (2, 2, 2, 2, 2))):
        # the TULIP constructor itself (tulip.py:531-584), as the BASELINE cfg5 surrogate needs it (SURVEY 8d option i)
        return TULIP(img_size=cfg.img_size, target_img_size=cfg.target_img_size, embed_dim=cfg.embed_dim, depths=cfg.depths,
                     num_heads=cfg.num_heads, mlp_ratio=4, qkv_bias=True, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.1,
                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **KW)
    fn = tulip_large if large else tulip_base
    return fn(img_size=cfg.img_size, target_img_size=cfg.target_img_size, **KW)


def load_params(model, pn):
    sd = {k: torch.from_numpy(v) for k, v in pn.items()}
    model.load_state_dict(sd, strict=True)                 # same strict load as misc.load_model (misc.py:382)


TULIP_WIDE = Cfg(embed_dim=192, depths=(2, 2, 18, 2), num_heads=(6, 12, 24, 48))      # BASELINE cfg5 surrogate (SURVEY 8d option i)
EXPANDING = Cfg(patch_unmerging=False)              # PatchExpanding instead of PatchUnmerging (tulip.py:126-141, 472, 565)
EXPANDING_HEAD = Cfg(patch_unmerging=False, pixel_shuffle=False)      # ... and FinalPatchExpanding instead of PixelShuffleHead (:144-159)
CASES = [("model_base_kitti_b2", TULIP_BASE, False), ("model_large_kitti_b1", TULIP_LARGE, True), ("model_expanding_kitti_b1", EXPANDING, False),
         ("model_expanding_head_kitti_b1", EXPANDING_HEAD, False),
         ("model_base_durlar_b1", Cfg(img_size=(32, 2048), target_img_size=(128, 2048)), False),
         ("model_wide_kitti_b1", TULIP_WIDE, False)]


@pytest.mark.parametrize("name,cfg,large", CASES)
def test_model_vs_reference_fixture(golden_dir, name, cfg, large):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    pn = make_params(cfg, int(g["pseed"]))
    lo, hi = make_inputs(cfg, int(g["batch"]), int(g["xseed"]))
    model = build(cfg, large).eval()
    load_params(model, pn)
    model.cuda()
    pred, loss, pixel = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda(), eval=True)
    st = int(g["pred_stride"])
    e_pred = rel_l2(pred[..., ::st], g["pred"])
    print(f"\n[{name}] pred rel-L2 {e_pred:.3e}  loss {loss.item():.6f} vs {float(g['loss']):.6f}  "
          f"pixel {pixel.item():.6f} vs {float(g['pixel_loss']):.6f}")
    assert pred.shape == (int(g["batch"]), 1, *cfg.target_img_size) and pred.dtype == torch.float32
    assert e_pred <= 1e-2
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    assert abs(pixel.item() - float(g["pixel_loss"])) <= 1e-3 * float(g["pixel_loss"])
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    grads = dict(model.named_parameters())
    ratios, heads = [], []
    for i, n in enumerate(names):
        gr = grads[n].grad
        assert gr is not None and gr.dtype == torch.float32 and gr.shape == grads[n].shape, n
        ratios.append(gr.double().norm().item() / max(float(g["grad_norm"][i]), 1e-30))
        k = min(64, gr.numel())
        heads.append(rel_l2(gr.reshape(-1)[:k], g["grad_head"][i][:k]))
    ratios = np.array(ratios)
    vec = np.array([grads[n].dim() == 1 for n in names])
    print(f"[{name}] grad-norm ratio min {ratios.min():.3f} max {ratios.max():.3f}; head rel-L2 median {np.median(heads):.3e} "
          f"worst {np.max(heads):.3e} ({names[int(np.argmax(heads))]})")
    assert np.all(np.abs(ratios[~vec] - 1) <= 0.15) and np.all(np.abs(ratios[vec] - 1) <= 0.35)
    assert np.median(heads) <= 2e-2


@pytest.mark.parametrize("cfg,B", [(TULIP_BASE, 1), (EXPANDING, 2), (EXPANDING_HEAD, 2)], ids=["base", "expanding", "expanding_head"])
def test_model_vs_oracle_full_gradients(cfg, B):
    """Every gradient tensor in full against the oracle's autograd (tulip_base, KITTI shape): the shipped configuration and the
    PatchExpanding / FinalPatchExpanding variants (tulip.py:126-159)."""
    pn = make_params(cfg, 11)
    lo, hi = make_inputs(cfg, B, 12)
    p = O.to_torch(pn, requires_grad=True)
    pred_o, loss_o, pixel_o = O.forward(p, cfg, torch.from_numpy(lo), torch.from_numpy(hi), state={})
    loss_o.backward()
    model = build(cfg).eval()
    load_params(model, pn)
    model.cuda()
    pred, loss, pixel = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())
    (loss * 65536.0).backward()                           # GradScaler's initial scale arrives as grad_loss (misc.py:292-295)
    errs = {n: rel_l2(q.grad / 65536.0, p[n].grad) for n, q in model.named_parameters()}
    worst = max(errs, key=errs.get)
    med = float(np.median(list(errs.values())))
    print(f"\npred rel-L2 {rel_l2(pred, pred_o):.3e}; grad rel-L2 median {med:.3e}, worst {errs[worst]:.3e} ({worst})")
    assert rel_l2(pred, pred_o) <= 1e-2 and med <= 2e-2 and errs[worst] <= 0.3


@pytest.mark.parametrize("cfg", [TULIP_BASE, EXPANDING_HEAD], ids=["base", "expanding_head"])
def test_train_mode_droppath_and_state_dict_roundtrip(cfg):
    pn = make_params(cfg, 21)
    lo, hi = make_inputs(cfg, 4, 22)
    model = build(cfg)
    load_params(model, pn)
    model.cuda().train()
    nb = 14
    gen = torch.Generator().manual_seed(5)
    rates = torch.tensor(model._drop_rates()).repeat_interleave(2)
    keep = 1 - rates
    scales = (torch.floor(keep[:, None] + torch.rand(2 * nb, 4, generator=gen)) / keep[:, None]).float()
    scales[3, 1] = 0.0                                    # make sure a dropped sample is exercised
    pred, loss, _ = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda(), _drop_scales=scales.cuda())
    loss.backward()
    names = [f"layers.{s}.blocks.{b}" for s in range(4) for b in range(2)] + [f"layers_up.{u}.blocks.{b}" for u in range(3) for b in range(2)]
    ds = {n: (scales[2 * i], scales[2 * i + 1]) for i, n in enumerate(names)}
    p = O.to_torch(pn, requires_grad=True)
    pred_o, loss_o, _ = O.forward(p, cfg, torch.from_numpy(lo), torch.from_numpy(hi), state={}, drop_scales=ds)
    loss_o.backward()
    assert rel_l2(pred, pred_o) <= 1.5e-2
    errs = [rel_l2(q.grad, p[n].grad) for n, q in model.named_parameters()]
    assert np.median(errs) <= 2e-2
    # sampled masks: floor(keep + U) / keep takes only the values {0, 1/keep}
    s = model._sample_drop_scales(64, torch.device("cuda"))
    assert s.shape == (28, 64)
    for i in range(28):
        k = float(keep[i])
        assert set(np.round(s[i].cpu().numpy() * k, 5).tolist()) <= {0.0, 1.0}
    # state_dict round trip through a fresh module (checkpoint compatibility, misc.py:332-349,382)
    sd = {k: v.cpu().clone() for k, v in model.state_dict().items()}
    m2 = build(cfg).eval()
    m2.load_state_dict(sd, strict=True)
    m2.cuda()
    model.eval()
    with torch.no_grad():
        a = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())[0]
        b = m2(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())[0]
        c = m2(torch.from_numpy(lo).cuda(), None, mc_drop=True)           # MC-dropout call convention (engine:417-419)
    assert torch.equal(a, b) and torch.equal(b, c)


def test_full_size_batch32_properties():
    """BASELINE cfg2 size (B=32): size-independent properties instead of an oracle run --
    frames are independent, so a batch permutation permutes the outputs and per-frame results equal B=1 runs;
    gradients of a mean loss over a duplicated batch equal those of the single batch."""
    cfg = TULIP_BASE
    pn = make_params(cfg, 31)
    lo, hi = make_inputs(cfg, 32, 32)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg).eval()
    load_params(model, pn)
    model.cuda()
    pred, loss, pixel = model(lo_t, hi_t)
    assert torch.isfinite(pred).all() and torch.isfinite(loss)
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(0)).cuda()
    pred_p, loss_p, _ = model(lo_t[perm], hi_t[perm])
    assert torch.equal(pred_p, pred[perm])
    assert abs(loss_p.item() - loss.item()) <= 1e-5
    for b in (0, 13, 31):
        pb, lb, _ = model(lo_t[b:b + 1], hi_t[b:b + 1])
        assert torch.equal(pb, pred[b:b + 1])
    assert abs(loss.item() - (pred - hi_t).abs().mean().item()) <= 1e-5
    loss.backward()
    g32 = {n: q.grad.clone() for n, q in model.named_parameters()}
    model.zero_grad()
    _, l16, _ = model(lo_t[:16].repeat(2, 1, 1, 1), hi_t[:16].repeat(2, 1, 1, 1))
    l16.backward()
    g_dup = {n: q.grad.clone() for n, q in model.named_parameters()}
    model.zero_grad()
    _, l16b, _ = model(lo_t[:16], hi_t[:16])
    l16b.backward()
    errs = [rel_l2(g_dup[n], q.grad) for n, q in model.named_parameters()]
    assert max(errs) <= 1e-2, max(errs)                  # identical math up to atomic-add ordering of bf16-rounded partials
    assert all(torch.isfinite(v).all() for v in g32.values())


@pytest.mark.parametrize("name,cfg,large,B", [("durlar_b16", Cfg(img_size=(32, 2048), target_img_size=(128, 2048)), False, 16),
                                               ("large_kitti_b8", TULIP_LARGE, True, 8)])
def test_full_size_other_configs_properties(name, cfg, large, B):
    """BASELINE cfg3 (DurLAR 32x2048 -> 128x2048, batch 16) and the 5-stage tulip_large at batch 8, at full size: frames are
    independent, so selected frames must equal their B=1 runs bit for bit (different tile schedules, same arithmetic), the
    loss must equal the mean over pred, and a step must leave finite gradients on every parameter."""
    pn = make_params(cfg, 41)
    lo, hi = make_inputs(cfg, B, 42)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg, large).eval()
    load_params(model, pn)
    model.cuda()
    pred, loss, pixel = model(lo_t, hi_t)
    assert tuple(pred.shape) == (B, 1, *cfg.target_img_size) and torch.isfinite(pred).all() and torch.isfinite(loss)
    assert abs(loss.item() - (pred - hi_t).abs().mean().item()) <= 1e-5
    loss.backward()
    grads = {n: q.grad.clone() for n, q in model.named_parameters()}
    assert all(torch.isfinite(g).all() and g.abs().sum() > 0 for g in grads.values())
    model.zero_grad()
    for b in (0, B - 1):
        pb, _, _ = model(lo_t[b:b + 1], hi_t[b:b + 1])
        assert torch.equal(pb, pred[b:b + 1])
    # the same batch again, on CUDA-graph replay by now (third identical call): identical pred, gradients equal up to the
    # ordering of the fp32 atomics in the weight-gradient reductions
    for _ in range(2):
        model.zero_grad()
        pred2, loss2, _ = model(lo_t, hi_t)
        loss2.backward()
    assert torch.equal(pred2, pred)
    errs = [rel_l2(q.grad, grads[n]) for n, q in model.named_parameters()]
    assert max(errs) <= 1e-3, max(errs)


def oracle_chunked(cfg, pn, lo, hi, chunk):
    """Oracle forward + backward of the mean-L1 loss over the whole batch, run `chunk` frames at a time on the host (the fp32
    autograd tape of a full BASELINE batch would take 8-17 GB): losses and gradients of the chunks are averaged."""
    B = lo.shape[0]
    p = O.to_torch(pn, requires_grad=True)
    preds, loss = [], 0.0
    for b0 in range(0, B, chunk):
        pr, l, _ = O.forward(p, cfg, torch.from_numpy(lo[b0:b0 + chunk]), torch.from_numpy(hi[b0:b0 + chunk]), state={})
        (l * (pr.shape[0] / B)).backward()
        preds.append(pr.detach())
        loss += l.item() * pr.shape[0] / B
    return torch.cat(preds), loss, {k: v.grad for k, v in p.items() if v.grad is not None}


@pytest.mark.parametrize("name,cfg,B,chunk", [("kitti_b32", TULIP_BASE, 32, 8),
                                               ("durlar_b16", Cfg(img_size=(32, 2048), target_img_size=(128, 2048)), 16, 4),
                                               ("wide_b8", TULIP_WIDE, 8, 2)])
def test_full_size_vs_oracle(name, cfg, B, chunk):
    """BASELINE cfg2 / cfg3 / cfg5-surrogate at their FULL batch sizes against the oracle on the same seeded inputs: pred, loss
    and every gradient tensor (these sizes were only covered by self-consistency properties)."""
    torch.set_num_threads(max(1, (os.cpu_count() or 1)))
    pn = make_params(cfg, 61)
    lo, hi = make_inputs(cfg, B, 62)
    pred_o, loss_o, grads_o = oracle_chunked(cfg, pn, lo, hi, chunk)
    model = build(cfg).eval()
    load_params(model, pn)
    model.cuda()
    pred, loss, _ = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())
    loss.backward()
    e_pred = rel_l2(pred, pred_o)
    errs = {n: rel_l2(q.grad, grads_o[n]) for n, q in model.named_parameters()}
    worst = max(errs, key=errs.get)
    med = float(np.median(list(errs.values())))
    print(f"\n[{name}] pred rel-L2 {e_pred:.3e}  loss {loss.item():.6f} vs {loss_o:.6f}  grad rel-L2 median {med:.3e} "
          f"worst {errs[worst]:.3e} ({worst})")
    assert e_pred <= 1e-2 and abs(loss.item() - loss_o) <= 1e-3 * loss_o
    assert med <= 2e-2 and errs[worst] <= 0.3


def test_gradient_accumulation_and_buffer_aliasing():
    """.grad tensors are views of one flat buffer; a second backward without zero_grad must accumulate, not alias."""
    from tulip_b200.parallel import flat_grad_of
    cfg = TULIP_BASE
    pn = make_params(cfg, 41)
    lo, hi = make_inputs(cfg, 2, 42)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg).eval()
    load_params(model, pn)
    model.cuda()
    model(lo_t, hi_t)[1].backward()
    flat = flat_grad_of(model)
    g1 = flat.clone()
    model(lo_t, hi_t)[1].backward()                       # accumulate
    assert rel_l2(flat_grad_of(model), 2 * g1) <= 1e-3
    model.zero_grad(set_to_none=True)
    model(lo_t, hi_t)[1].backward()
    assert rel_l2(flat_grad_of(model), g1) <= 1e-3
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    before = model._flat.clone()
    opt.step()
    assert not torch.equal(before, model._flat), "optimizer updates must land in the flat parameter buffer"


def test_reference_engine_call_pattern_trains():
    """The call sequence of the unchanged reference engine (engine_upsampling.py:77-100, util/misc.py:292-308): fp16 autocast
    context, GradScaler (initial scale 65536), unscale_, gradient-norm over p.grad, AdamW with no-decay groups, zero_grad.
    The loss must fall on a fixed batch and no inf/nan may be reported to the scaler."""
    cfg = TULIP_BASE
    torch.manual_seed(0)
    model = build(cfg).cuda().train()
    lo, hi = make_inputs(cfg, 4, 52)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    decay = [p for n, p in model.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in model.named_parameters() if p.ndim <= 1]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=5e-4, betas=(0.9, 0.95))
    scaler = torch.amp.GradScaler("cuda")
    losses = []
    for it in range(12):
        with torch.autocast("cuda"):
            _, total_loss, pixel_loss = model(lo_t, hi_t, eval=False)
        losses.append(total_loss.item())
        assert np.isfinite(losses[-1]) and np.isfinite(pixel_loss.item())
        total_loss /= 1                                       # engine_upsampling.py:92 (`total_loss /= accum_iter`, in place)
        scaler.scale(total_loss).backward()
        scaler.unscale_(opt)
        norm = torch.norm(torch.stack([torch.norm(p.grad.detach(), 2.0) for p in model.parameters()]), 2.0)
        assert torch.isfinite(norm)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        torch.cuda.synchronize()
    assert scaler.get_scale() == 65536.0, "a step was skipped: inf/nan gradients"
    assert losses[-1] < 0.8 * losses[0], losses


def test_deepcopy_and_pickle_after_forward():
    """EMA wrappers deep-copy the model (the reference's train_one_epoch has an `ema` hook); after the first forward
    the module holds a C handle and raw pointers, which must not travel with the copy."""
    import copy
    import io
    cfg = TULIP_BASE
    model = build(cfg).cuda().eval()
    lo, hi = make_inputs(cfg, 1, 7)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    pred, _, _ = model(lo_t, hi_t)
    twin = copy.deepcopy(model)
    assert twin._net is None
    pred2, _, _ = twin(lo_t, hi_t)
    assert twin._net is not None and twin._net.value != model._net.value
    assert torch.equal(pred, pred2)
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert torch.equal(again(lo_t, hi_t)[0], pred)


def test_parameter_storage_swap_is_noticed():
    """replacing a parameter's storage (EMA swap, load_state_dict(assign=True)) must reach the next forward."""
    cfg = TULIP_BASE
    model = build(cfg).cuda().eval()
    lo, hi = make_inputs(cfg, 1, 7)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    pred, _, _ = model(lo_t, hi_t)
    p = model.layers[1].blocks[0].mlp.fc1.weight                # not one of the spot-checked parameters of round 1
    p.data = p.data.clone() * 1.5
    pred2, _, _ = model(lo_t, hi_t)
    assert not torch.equal(pred, pred2)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["norm_up.weight"] = sd["norm_up.weight"] * 2.0
    model.load_state_dict(sd, assign=True)
    pred3, _, _ = model(lo_t, hi_t)
    assert not torch.equal(pred2, pred3)


@pytest.mark.parametrize("cfg,large", [(TULIP_BASE, False), (TULIP_LARGE, True)])
def test_backward_in_three_phases_equals_one_pass(cfg, large):
    """tulip_net_backward_phases (the host of the overlapped gradient all-reduce): phases 0, 1, 2 issued as three calls give the
    gradients of the single call -- bit for bit where no atomics are involved, to fp32 atomic-order noise elsewhere."""
    pn = make_params(cfg, 71)
    lo, hi = make_inputs(cfg, 2, 72)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg, large).eval()
    load_params(model, pn)
    model.cuda()
    _, loss, _ = model(lo_t, hi_t)
    loss.backward()
    want = {n: q.grad.clone() for n, q in model.named_parameters()}
    model._force_phases = True
    for _ in range(3):                                        # third call replays the three per-phase CUDA graphs
        model.zero_grad()
        _, loss, _ = model(lo_t, hi_t)
        loss.backward()
    errs = {n: rel_l2(q.grad, want[n]) for n, q in model.named_parameters()}
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1e-4, (worst, errs[worst])


def test_predrawn_droppath_scales_are_reproducible_and_invalidate():
    """The DropPath scales of the next training forward are drawn right after the current forward is launched (off the host's
    critical path when the loop reads the loss every step).  Same seed -> same sequence of scale sets; a reseed or any other
    CUDA random op in between discards the pre-drawn set instead of using stale numbers."""
    cfg = TULIP_BASE
    pn = make_params(cfg, 41)
    lo, hi = make_inputs(cfg, 2, 42)
    lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    model = build(cfg)
    load_params(model, pn)
    model.cuda().train()

    def run(n, disturb=False):
        torch.manual_seed(7)
        out = []
        for i in range(n):
            if disturb and i == 1:
                torch.rand(3, device="cuda")                  # moves the generator: the pre-drawn set must not be used
            _, loss, _ = model(lo_t, hi_t)
            loss.item()
            (bufs,) = [b for b in model._step_bufs.values()]
            out.append(bufs["drop"].clone())                  # the scales this forward ran with
        return out

    a = run(8)
    hits = model._predraw_hits
    assert hits == 7 and "_predrawn" in model.__dict__        # every call after the first used the set drawn behind its predecessor
    b = run(8)
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and model._predraw_hits == hits + 7   # reseed: waiting set discarded, same replay
    c = run(8, disturb=True)
    assert torch.equal(c[0], a[0]) and not all(torch.equal(x, y) for x, y in zip(a[1:], c[1:]))   # the stream moved: nothing stale reused
    assert model._predraw_hits == hits + 7 + 6
    assert not all(torch.equal(a[0], x) for x in a[1:])       # masks do change from step to step


def test_stage_inputs_one_launch_copy():
    from tulip_b200._lib import check, current_stream, load_library, ptr
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(0)
    for sizes in ((32 * 16 * 1024, 32 * 64 * 1024, 28 * 32), (1027, 0, 5), (4, 8, 0)):
        src = [torch.rand(n, device="cuda", generator=g) if n else None for n in sizes]
        dst = [torch.full((n + 4,), -1.0, device="cuda") if n else None for n in sizes]
        check(lib.tulip_stage_inputs(ptr(src[0]), ptr(dst[0]), sizes[0], ptr(src[1]), ptr(dst[1]), sizes[1], ptr(src[2]), ptr(dst[2]),
                                     sizes[2], current_stream()), "tulip_stage_inputs")
        for s_, d_, n in zip(src, dst, sizes):
            if n:
                assert torch.equal(d_[:n], s_) and bool((d_[n:] == -1).all())


def test_loss_item_reads_behind_the_forward_only():
    """TULIP.loss_item(): the pinned-host copy of (total, pixel) queued behind the forward equals `.item()` on the returned
    tensors, step after step on the persistent buffers, also when a backward pass is queued before the read."""
    cfg = TULIP_BASE
    torch.manual_seed(0)
    model = build(cfg).cuda().train()
    with pytest.raises(RuntimeError):
        model.loss_item()
    for it in range(4):
        lo, hi = make_inputs(cfg, 2, 60 + it)
        _, total_loss, pixel_loss = model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())
        total_loss.backward()
        got, got_pixel = model.loss_item(), model.loss_item(pixel=True)
        assert got == total_loss.item() and got_pixel == pixel_loss.item()
        model.zero_grad(set_to_none=True)
    with torch.no_grad():                                      # a forward-only call in between invalidates the read-back
        model(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())
    with pytest.raises(RuntimeError):
        model.loss_item()
    import copy
    clone = copy.deepcopy(model)                               # streams / events are runtime state, not module state
    assert "_loss_rb" not in clone.__dict__
