"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
its parameter schema equals the oracle's restatement of the reference state_dict, and the host-side module
mirrors the reference constructor contract.  No kernels are launched."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle.params import TULIP_BASE, TULIP_LARGE, Cfg, param_shapes
from tulip_b200._lib import SIGNATURES, TulipConfig, load_library
from tulip_b200.model.tulip import TULIP, WindowAttention, tulip_base, tulip_large

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(img_size=(16, 1024), target_img_size=(64, 1024), patch_size=(1, 4), in_chans=1, window_size=[2, 8], swin_v2=False,
          pixel_shuffle=True, circular_padding=True, log_transform=True, patch_unmerging=True)


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tulip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tulip_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libtulip_b200.so does not export {s}"
    assert sorted(SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert lib.tulip_abi_version() == 2


def c_schema(cfg: Cfg):
    lib = load_library()
    c = TulipConfig()
    c.img_h, c.img_w = cfg.img_size
    c.tgt_h, c.tgt_w = cfg.target_img_size
    c.patch_h, c.patch_w = cfg.patch_size
    c.in_chans, c.embed_dim = cfg.in_chans, cfg.embed_dim
    c.win_h, c.win_w = cfg.window_size
    c.num_layers = cfg.num_layers
    for i in range(cfg.num_layers):
        c.depths[i], c.num_heads[i] = cfg.depths[i], cfg.num_heads[i]
    c.mlp_ratio, c.ln_eps, c.log_transform = cfg.mlp_ratio, cfg.ln_eps, int(cfg.log_transform)
    c.patch_expanding, c.expanding_head = int(not cfg.patch_unmerging), int(not cfg.pixel_shuffle)
    h = C.c_void_p()
    rc = lib.tulip_net_create(C.byref(c), C.byref(h))
    if rc != 0:
        return rc, lib.tulip_last_error().decode()
    out = []
    buf, shape, nd = C.create_string_buffer(256), (C.c_int64 * 4)(), C.c_int()
    for i in range(lib.tulip_net_num_params(h)):
        assert lib.tulip_net_param_info(h, i, buf, 256, shape, C.byref(nd)) == 0
        out.append((buf.value.decode(), tuple(shape[:nd.value])))
    ws = lib.tulip_net_workspace_bytes(h, 2)
    nblk = lib.tulip_net_num_blocks(h)
    lib.tulip_net_destroy(h)
    return out, ws, nblk


@pytest.mark.parametrize("cfg,nblk", [(TULIP_BASE, 14), (TULIP_LARGE, 18),
                                      (Cfg(img_size=(32, 2048), target_img_size=(128, 2048)), 14),
                                      (Cfg(patch_unmerging=False), 14),           # PatchExpanding variant (tulip.py:126-141)
                                      (Cfg(patch_unmerging=False, pixel_shuffle=False), 14)])   # + FinalPatchExpanding (:144-159)
def test_c_schema_matches_reference_state_dict(cfg, nblk):
    schema, ws, n = c_schema(cfg)
    want = [(k, tuple(v)) for k, v in param_shapes(cfg).items() if not k.endswith("relative_position_index")]
    assert schema == want
    assert n == nblk and ws > 0


def test_c_config_errors_are_reported_not_fatal():
    bad = Cfg(window_size=(4, 8))
    rc, msg = c_schema(bad)
    assert rc != 0 and "16 tokens" in msg
    rc, msg = c_schema(Cfg(target_img_size=(128, 1024)))          # BASELINE cfg5's 8x head: the reference cannot build it either
    assert rc != 0 and "upscale_factor" in msg
    rc, msg = c_schema(Cfg(pixel_shuffle=False, embed_dim=192, num_heads=(6, 12, 24, 48)))   # FinalPatchExpanding: embed_dim 96 only
    assert rc != 0 and "FinalPatchExpanding" in msg


@pytest.mark.parametrize("factory,cfg", [(tulip_base, TULIP_BASE), (tulip_large, TULIP_LARGE), (tulip_base, Cfg(patch_unmerging=False)),
                                         (tulip_base, Cfg(patch_unmerging=False, pixel_shuffle=False))])
def test_module_state_dict_schema(factory, cfg):
    m = factory(**{**KW, "patch_unmerging": cfg.patch_unmerging, "pixel_shuffle": cfg.pixel_shuffle})
    sd = m.state_dict()
    want = param_shapes(cfg)
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k])
        assert v.dtype == (torch.int64 if k.endswith("relative_position_index") else torch.float32)
    assert [n for n, _ in m.named_children()] == ["pos_drop", "layers", "layers_up", "first_patch_expanding",
                                                  "skip_connection_layers", "norm_up", "patch_embed", "decoder_pred",
                                                  "ps_head" if cfg.pixel_shuffle else "final_patch_expanding"]
    assert m.upscale_factor == 4
    m._create_net()                      # C schema check against this module (raises on mismatch)


def test_module_contract_details():
    m = tulip_base(**KW)
    rates = m._drop_rates()
    want = [0, .1 / 7, .2 / 7, .3 / 7, .4 / 7, .5 / 7, .6 / 7, .1, .4 / 7, .5 / 7, .2 / 7, .3 / 7, 0, .1 / 7]
    assert np.allclose(rates, want, atol=1e-6)
    assert isinstance(m.layers[0].blocks[0].drop_path, torch.nn.Identity)
    a = WindowAttention(96, [2, 8], 3, shift=True)
    assert a.shift_size == (1, 4) and a.scale == 32 ** -0.5
    assert a.window_mode(2) == 0 and a.window_size == (2, 8)
    assert a.window_mode(1) == 1 and a.window_size == (1, 16) and a.shift_size == (0, 8)
    assert a.window_mode(4) == 1, "the backup switch is permanent (tulip.py:284-287)"
    with pytest.raises(NotImplementedError):
        tulip_base(**{**KW, "swin_v2": True})
    with pytest.raises(NotImplementedError):
        tulip_base(**{**KW, "circular_padding": False})
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 16, 1024), torch.zeros(1, 1, 64, 1024))


@pytest.mark.skipif(not os.path.exists("/root/reference/tulip/model/tulip.py"), reason="reference tree only exists in the build container")
def test_init_matches_reference_rng_stream():
    """Same seed -> bit-identical initial state_dict as the reference (construction order and init calls match)."""
    from oracle.make_golden import build_reference, import_reference
    T = import_reference()
    torch.manual_seed(0)
    ref = build_reference(T, TULIP_BASE, False).state_dict()
    torch.manual_seed(0)
    ours = tulip_base(**KW).state_dict()
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k


def test_gemm_tiling_plan_host_logic():
    """The tile-width / schedule decision of the tcgen05 GEMM launcher is pure host logic (tulip_gemm_nt_plan): check its
    invariants for every Linear shape of the KITTI / DurLAR steps at the bench batch sizes, without a GPU."""
    import ctypes as C
    from tulip_b200._lib import load_library
    lib = load_library()
    keys = ("bn", "panel", "n_chunks", "npc", "nworkers", "nsa", "kb", "klast", "grid", "stages")
    seen_panel = seen_pairs = 0
    prev_pairs = lib.tulip_gemm_nt_pairs_mode(2)               # the CTA-pair schedule is opt-in: plan it for K >= 384 here
    for tokens0, in ((32 * 16 * 256,), (16 * 32 * 512,)):
        for s in range(4):
            T, Cc = tokens0 >> (2 * s), 96 << s
            for (N, K, epi) in ((3 * Cc, Cc, 0), (Cc, Cc, 2), (4 * Cc, Cc, 1), (Cc, 4 * Cc, 2), (Cc, 3 * Cc, 0), (Cc, 4 * Cc, 0), (4 * Cc, Cc, 5)):
                out = (C.c_int * 10)()
                assert lib.tulip_gemm_nt_plan(T, N, K, epi, 0, out) == 0
                p = dict(zip(keys, out))
                assert p["bn"] in (96, 192) and N % p["bn"] == 0
                assert p["kb"] == -(-K // 64) and 1 <= p["klast"] <= 4 and 16 * (p["klast"] - 1) < K - 64 * (p["kb"] - 1) <= 16 * p["klast"]
                assert 1 <= p["grid"] <= 148 and 2 <= p["stages"] <= 8
                if p["panel"] == 2:                                                                # CTA pairs: deep K, whole pairs
                    seen_pairs += 1
                    assert K >= 384 and p["grid"] % 2 == 0 and epi in (0, 1, 2)
                if p["panel"] == 1:
                    seen_panel += 1
                    assert p["n_chunks"] * p["npc"] * p["bn"] == N and p["grid"] == p["n_chunks"] * p["nworkers"]
                    assert p["nsa"] >= (p["kb"] + 1 if p["npc"] > 1 else 3) and p["nsa"] <= 8
                    area = p["stages"] * (128 * 64 * 2 + p["bn"] * 64 * 2)
                    assert p["npc"] * p["kb"] * p["bn"] * 128 + p["nsa"] * 128 * 64 * 2 <= area      # resident B + A ring fit the ring area
                    assert -(-T // 128) >= 4 * p["nworkers"]                                       # enough row panels per worker
    assert seen_panel >= 8                                                                         # the wide stages do use it
    assert seen_pairs >= 4                                                                         # the deep-K GEMMs of stages 2-3 run as pairs
    lib.tulip_gemm_nt_pairs_mode(max(prev_pairs, 0))
    out = (C.c_int * 10)()
    assert lib.tulip_gemm_nt_plan(2048, 768, 3072, 2, 0, out) == 0 and out[1] == (2 if prev_pairs > 0 else 0)   # default: no pairs
    out = (C.c_int * 10)()
    assert lib.tulip_gemm_nt_plan(1000, 100, 96, 0, 0, out) != 0                                   # N % 96: not a tcgen05 shape
    assert lib.tulip_gemm_nt_plan(131072, 384, 96, 1, 1, out) == 0 and out[0] == 96              # saved pre-activation: narrow tile


def test_tn_group_plan_host_logic():
    """Work-item cut of the grouped weight-gradient launch (tulip_gemm_tn_group_plan, pure host logic): every token block of
    every problem is covered exactly once, ranges keep >= 4 blocks where the problem has them, and a round-robin deal over
    148 SMs (the kernel hands items out dynamically; this is the planner's model) is balanced for the half-block groups of
    BASELINE cfg2 (B = 32)."""
    from tulip_b200 import ops
    for T, C in [(131072, 96), (32768, 192), (8192, 384), (2048, 768)]:
        for shapes in ([(T, C, 4 * C), (T, 4 * C, C)], [(T, C, C), (T, 3 * C, C)]):
            per, items = ops.gemm_tn_group_plan(shapes, sms=148)
            loads = [0.0] * 148
            idx = 0
            for (M, N, K), pp in zip(shapes, per):
                tb = -(-M // 64)
                assert 1 <= pp <= tb and (pp >= 4 or pp == tb)
                splits = -(-tb // pp)
                assert (splits - 1) * pp < tb
                for s_ in range(splits):
                    ntb = min(tb, (s_ + 1) * pp) - s_ * pp
                    for nt in range(-(-N // 128)):
                        for kt in range(-(-K // 192)):
                            loads[idx % min(items, 148)] += ntb * (2 + -(-min(192, K - kt * 192) // 64))
                            idx += 1
            assert idx == items
            used = [v for v in loads if v > 0]
            assert len(used) >= 100                                   # one wave may be smaller than the chip (an item pays a fixed
            assert max(used) <= 1.25 * (sum(used) / len(used))         # cost), but the longest SM stays within 25 % of the mean
    with pytest.raises(RuntimeError):
        ops.gemm_tn_group_plan([(0, 96, 96)])
