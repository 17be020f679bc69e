"""ctypes binding of libtulip_b200.so (C ABI: include/tulip_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_STAGES = 8


class TulipLibraryError(RuntimeError):
    pass


class TulipConfig(C.Structure):
    _fields_ = [
        ("img_h", C.c_int), ("img_w", C.c_int), ("tgt_h", C.c_int), ("tgt_w", C.c_int),
        ("patch_h", C.c_int), ("patch_w", C.c_int), ("in_chans", C.c_int), ("embed_dim", C.c_int),
        ("win_h", C.c_int), ("win_w", C.c_int), ("num_layers", C.c_int),
        ("depths", C.c_int * MAX_STAGES), ("num_heads", C.c_int * MAX_STAGES),
        ("mlp_ratio", C.c_int), ("ln_eps", C.c_float), ("log_transform", C.c_int),
        ("patch_expanding", C.c_int), ("expanding_head", C.c_int),
    ]


class GemmDesc(C.Structure):                     # tulip_gemm_desc (include/tulip_b200.h)
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("A2", C.c_void_p), ("lda2", C.c_int64), ("K1", C.c_int),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("B2", C.c_void_p), ("ldb2", C.c_int64), ("K2", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("a_mode", C.c_int), ("g_H", C.c_int), ("g_W", C.c_int), ("g_Cc", C.c_int),
                ("bias", C.c_void_p),
                ("out", C.c_void_p), ("ldo", C.c_int64), ("out2", C.c_void_p), ("ldo2", C.c_int64),
                ("aux", C.c_void_p), ("ldaux", C.c_int64),
                ("row_scale", C.c_void_p), ("rows_per_sample", C.c_int), ("split_col", C.c_int),
                ("wd", C.c_void_p), ("target", C.c_void_p), ("pred", C.c_void_p), ("gscale", C.c_void_p), ("dwd", C.c_void_p),
                ("hd_H", C.c_int), ("hd_W", C.c_int), ("hd_r", C.c_int), ("hd_E", C.c_int),
                ("aux2", C.c_void_p), ("ldaux2", C.c_int64),
                ("ln_w", C.c_void_p), ("ln_stats", C.c_void_p), ("ln_dw", C.c_void_p), ("ln_db", C.c_void_p),
                ("ln_b", C.c_void_p), ("ln_y", C.c_void_p), ("ln_ystats", C.c_void_p), ("ln_eps", C.c_float), ("hd_ln", C.c_int)]


class GemmTNDesc(C.Structure):                   # tulip_gemm_tn_desc
    _fields_ = [("dY", C.c_void_p), ("ldy", C.c_int64), ("X", C.c_void_p), ("ldx", C.c_int64), ("X2", C.c_void_p), ("ldx2", C.c_int64),
                ("K1", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("y_mode", C.c_int), ("g_H", C.c_int), ("g_W", C.c_int), ("g_Cc", C.c_int),
                ("dW", C.c_void_p), ("lddw", C.c_int64), ("db", C.c_void_p), ("perm_R2", C.c_int), ("perm_Cc", C.c_int)]


def lib_path() -> str:
    return os.environ.get("TULIP_B200_LIB", os.path.join(_HERE, "lib", "libtulip_b200.so"))


_vp, _fp, _i, _i64 = C.c_void_p, C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes); must list every symbol include/tulip_b200.h declares
SIGNATURES = {
    "tulip_last_error": (C.c_char_p, []),
    "tulip_abi_version": (_i, []),
    "tulip_net_create": (_i, [C.POINTER(TulipConfig), C.POINTER(_vp)]),
    "tulip_net_destroy": (None, [_vp]),
    "tulip_net_num_params": (_i, [_vp]),
    "tulip_net_param_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i64), C.POINTER(_i)]),
    "tulip_net_num_blocks": (_i, [_vp]),
    "tulip_net_block_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tulip_net_workspace_bytes": (_i64, [_vp, _i]),
    "tulip_net_kernel_launches": (_i64, [_vp]),
    "tulip_net_profile": (_i, [_vp, _i]),
    "tulip_net_profile_num_tags": (_i, []),
    "tulip_net_profile_read": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(_i64)]),
    "tulip_net_profile_record": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tulip_net_set_inference": (_i, [_vp, _i]),
    "tulip_net_set_params_version": (_i, [_vp, C.c_longlong]),
    "tulip_net_profile_where": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tulip_net_forward": (_i, [_vp, _i, _fp, _vp, _fp, _fp, _fp, _vp, _vp, _fp, _fp, _vp]),
    "tulip_net_backward": (_i, [_vp, _i, _fp, _vp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _vp, _vp]),
    "tulip_net_backward_phases": (_i, [_vp, _i, _fp, _vp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _vp, _vp, _i, _i]),
    "tulip_gemm_nt": (_i, [_vp, _vp, _fp, _vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "tulip_gemm_nt_plan": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i)]),
    "tulip_gemm_nt_pairs_mode": (_i, [_i]),
    "tulip_set_sm_budget": (_i, [_i]),
    "tulip_gemm_tn": (_i, [_vp, _vp, _fp, _fp, _i, _i, _i, _i, _vp]),
    "tulip_gemm_nt_ex": (_i, [C.POINTER(GemmDesc), _i, _vp]),
    "tulip_gemm_tn_ex": (_i, [C.POINTER(GemmTNDesc), _vp]),
    "tulip_gemm_tn_group": (_i, [C.POINTER(GemmTNDesc), _i, _vp]),
    "tulip_gemm_tn_group_plan": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "tulip_head_bwd_fused_supported": (_i, [_i, _i]),
    "tulip_head_bwd_fused": (_i, [_vp, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _vp]),
    "tulip_wmsa_block_supported": (_i, [_i] * 7),
    "tulip_wmsa_block_fwd": (_i, [_vp, _vp, _fp, _fp, _vp, _fp, _vp, _fp, _fp, _fp] + [_i] * 12 + [C.c_float, _vp]),
    "tulip_mlp_block_supported": (_i, [_i, _i]),
    "tulip_mlp_block_fwd": (_i, [_vp, _vp, _fp, _fp, _vp, _fp, _vp, _fp, _fp, _i, _vp, _fp, _vp, _i, _i, C.c_float, _vp]),
    "tulip_window_attention_fwd": (_i, [_vp, _fp, _vp] + [_i] * 12 + [_vp]),
    "tulip_window_attention_bwd": (_i, [_vp, _fp, _vp, _vp, _fp] + [_i] * 12 + [_vp]),
    "tulip_layernorm_fwd": (_i, [_vp, _fp, _fp, _vp, _fp, _i, _i, C.c_float, _i, _i, _i, _vp]),
    "tulip_layernorm_bwd": (_i, [_vp, _fp, _fp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "tulip_patch_embed_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, C.c_float, _vp]),
    "tulip_patch_embed_bwd": (_i, [_fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, C.c_float, _vp]),
    "tulip_eval_postprocess": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, C.c_float, _i, _vp]),
    "tulip_mc_dropout_aggregate": (_i, [_fp, _fp, _fp, _i, _i64, C.c_float, _vp]),
    "tulip_range_to_points": (_i, [_fp, _fp, _fp, _fp, _fp, C.c_float, _fp, _i, _i, _i, _vp]),
    "tulip_voxel_metrics_workspace_bytes": (_i64, [_i]),
    "tulip_voxel_metrics": (_i, [_fp, _fp, _i, C.c_float, _vp, _vp, _vp]),
    "tulip_range_to_points_durlar": (_i, [_fp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_double, C.c_double, _vp, _i, _i, _i, _vp]),
    "tulip_voxel_metrics_f64": (_i, [_vp, _vp, _i, C.c_double, _vp, _vp, _vp]),
    "tulip_chamfer_distance": (_i, [_fp, _fp, _i, _i, _fp, _fp, _fp, _vp]),
    "tulip_preprocess_range": (_i, [_fp, _i, C.c_float, _i, C.c_float, C.c_float, _i, _i, _i, _fp, _fp, _i, _i, _i, _vp]),
    "tulip_rimg_decode": (_i, [_vp, _fp, _i, _i, _i, _vp]),
    "tulip_adamw_step": (_i, [_fp, _fp, _fp, _fp, _vp, _i, _i64, _vp, _vp]),
    "tulip_grad_norm": (_i, [_fp, _vp, _i, _i64, _vp, _fp, _vp]),
    "tulip_l1_loss": (_i, [_fp, _fp, _i64, _i, _fp, _fp, _vp]),
    "tulip_stage_inputs": (_i, [_fp, _fp, _i64, _fp, _fp, _i64, _fp, _fp, _i64, _vp]),
    "tulip_window_partition": (_i, [_vp, _vp] + [_i] * 8 + [_vp]),
    "tulip_window_reverse": (_i, [_vp, _vp] + [_i] * 8 + [_vp]),
    "tulip_shift_mask": (_i, [_fp] + [_i] * 6 + [_vp]),
    "tulip_rel_bias_gather": (_i, [_fp, _fp, _i, _i, _i, _vp]),
    "tulip_merge_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tulip_pixel_shuffle": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
}


def load_library():
    """dlopen the CUDA library; raises TulipLibraryError (never falls back) if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise TulipLibraryError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tulip_b200/csrc`). tulip_b200 has no CPU / PyTorch fallback.")
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise TulipLibraryError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TulipLibraryError(f"{path} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = "tulip_b200"):
    if rc != 0:
        msg = load_library().tulip_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
