"""ctypes binding of libmccnn_b200.so (include/mccnn_b200.h).

The shared library is the product; this module only loads it and declares the
signatures. There is no CPU fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmccnn_b200.so")


class SgmParams(C.Structure):
    _fields_ = [("P1", C.c_float), ("P2", C.c_float), ("P1_red", C.c_float), ("P2_red", C.c_float),
                ("threshold", C.c_int), ("subpixel", C.c_int), ("bilateral", C.c_int),
                ("cbca_iters", C.c_int), ("cbca_L1", C.c_int), ("cbca_tau", C.c_int)]


class Shard(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("H_full", C.c_int), ("row0", C.c_int), ("rows", C.c_int),
                ("xchg_local", C.c_void_p), ("xchg_prev", C.c_void_p), ("xchg_next", C.c_void_p), ("epoch", C.c_uint),
                ("go_flag", C.c_void_p), ("timeout_ms", C.c_uint)]


class FcWeights(C.Structure):
    """mccnn_fc_weights: device pointers of the MC-CNN-accurate head (include/mccnn_b200.h)."""
    _fields_ = [("w1_left", C.c_void_p), ("w1_right", C.c_void_p), ("b1", C.c_void_p), ("w2_blocks_f16", C.c_void_p),
                ("b2", C.c_void_p), ("w3_blocks_f16", C.c_void_p), ("b3", C.c_void_p), ("w4", C.c_void_p), ("b4", C.c_float)]


_vp, _sz, _i, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_float
_PP = C.POINTER(SgmParams)

# name -> (restype, argtypes); must list every symbol include/mccnn_b200.h declares
SIGNATURES = {
    "mccnn_last_error": (C.c_char_p, []),
    "mccnn_abi_version": (_i, []),
    "mccnn_default_sgm_params": (None, [_PP]),
    "mccnn_disp_pitch": (_i, [_i]),
    "mccnn_device_supported": (_i, [_i]),
    "mccnn_standardize_pad": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mccnn_pad_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_conv_packed_weight_bytes": (_sz, [_i]),
    "mccnn_pack_weights_host": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _vp]),
    "mccnn_conv_workspace_bytes": (_sz, [_i, _i, _i]),
    "mccnn_conv_tower": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "mccnn_conv_tower_fp32": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "mccnn_cost_volume": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "mccnn_cost_volume_fast": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "mccnn_cost_volume_fast_tc_workspace_bytes": (_sz, [_i, _i]),
    "mccnn_cost_volume_fast_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp]),
    "mccnn_cost_volume_tc_workspace_bytes": (_sz, [_i, _i]),
    "mccnn_cost_volume_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp]),
    "mccnn_fc_matrix_blocks_bytes": (_sz, []),
    "mccnn_pack_fc_matrix_host": (_i, [_vp, _vp]),
    "mccnn_fc_head_workspace_bytes": (_sz, [_i, _i]),
    "mccnn_cost_volume_accurate": (_i, [_vp, _vp, C.POINTER(FcWeights), _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp]),
    "mccnn_volume_to_dhw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_cross_arms": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "mccnn_cbca": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mccnn_sgm_workspace_bytes": (_sz, [_i, _i, _i]),
    "mccnn_sgm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _PP, _i, _i, _vp]),
    "mccnn_sgm_shard_exchange_bytes": (_sz, [_i]),
    "mccnn_sgm_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _PP, _i, _i, C.POINTER(Shard), _i, _vp]),
    "mccnn_sgm_fused_shard_exchange_bytes": (_sz, [_i, _i]),
    "mccnn_sgm_fused_sharded": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _PP, _i, C.POINTER(Shard), _i, _vp]),
    "mccnn_sgm_shard_status": (_i, [_vp, C.POINTER(_i), _vp]),
    "mccnn_sgm_single_path": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _PP, _i, _vp]),
    "mccnn_wta": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_wta_dhw": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_lr_flags": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "mccnn_lrc_fill_workspace_bytes": (_sz, [_i, _i]),
    "mccnn_lrc_fill": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "mccnn_median5": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mccnn_bilateral9": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mccnn_encode_u8": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_encode_u16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_wta_subpixel": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mccnn_bad_pixels": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mccnn_bad_pixels_u16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mccnn_pipeline_workspace_bytes": (_sz, [_i, _i, _i]),
    "mccnn_disparity_pipeline": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _PP, _i, _vp, _vp]),
    "mccnn_match_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mccnn_match_pair": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _PP, _i, _vp, _vp]),
    "mccnn_train_param_count": (_sz, [_i]),
    "mccnn_train_workspace_bytes": (_sz, [_i, _i, _i]),
    "mccnn_train_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _f, _f, _i, _vp]),
    "mccnn_match_accurate_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mccnn_match_pair_accurate": (_i, [_vp, _vp, _vp, C.POINTER(FcWeights), _vp, _vp, _vp, _sz, _i, _i, _i, _i, _PP, _i, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C scenedepthestimation_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "mccnn") -> None:
    if rc != 0:
        msg = load().mccnn_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode(errors='replace') if msg else ''}")


def default_sgm_params() -> SgmParams:
    p = SgmParams()
    load().mccnn_default_sgm_params(C.byref(p))
    return p
