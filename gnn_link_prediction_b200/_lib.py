"""ctypes binding of libhgin.so (include/hgin.h) — the only way the host code reaches the GPU
kernels.  There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgin.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hgin.h")

OK = 0
SELF_NONE, SELF_ADD, SELF_CONCAT = 0, 1, 2
ACT_NONE, ACT_PRELU, ACT_RELU = 0, 1, 2
ACT_LEAKY_RELU, ACT_ELU, ACT_SIGMOID, ACT_TANH, ACT_GELU, ACT_SILU, ACT_SOFTPLUS = 3, 4, 5, 6, 7, 8, 9
MATH_FP32, MATH_TF32, MATH_BF16 = 0, 1, 2
DTYPE_F32, DTYPE_BF16 = 0, 1

_i32, _i64, _f32, _f64, _ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p
_u64 = ctypes.c_uint64

class CollateField(ctypes.Structure):
    """hgin_collate_field (include/hgin.h)."""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("ptr", ctypes.c_void_p), ("width", ctypes.c_int32),
                ("size_class", ctypes.c_int32), ("add_class", ctypes.c_int32), ("closing_row", ctypes.c_int32)]


# name -> (restype, argtypes); must list every function include/hgin.h declares
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "hgin_version": (_i32, []),
    "hgin_last_error": (ctypes.c_char_p, []),
    "hgin_csr_workspace_bytes": (_i64, [_i64, _i64]),
    "hgin_csr_build": (_i32, [_ptr, _i32, _i64, _i64, _i32, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_gin_combine": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _i32, _i32, _ptr, _i64, _ptr]),
    "hgin_gin_combine_post_workspace_bytes": (_i64, []),
    "hgin_gin_combine_post": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _i32, _i32, _ptr, _i64,
                                     _ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_gin_combine_pre": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _i32, _i32, _ptr, _i64,
                                    _i32, _ptr, _i32, _ptr, _ptr]),
    "hgin_gin_combine_t": (_i32, [_i32, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _i32, _i32, _ptr,
                                  _i64, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_gin_combine_table_t": (_i32, [_i32, _i64, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _i64, _i32,
                                        _ptr, _i32, _i32, _ptr, _i64, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr,
                                        _ptr, _i64, _ptr]),
    "hgin_gin_combine_staged_t": (_i32, [_i32, _i64, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _i64, _ptr,
                                         _i32, _i32, _ptr, _i64, _i32, _ptr, _i32, _ptr, _ptr]),
    "hgin_block_gate": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr]),
    "hgin_gin_combine_blocks_t": (_i32, [_i32, _i64, _ptr, _ptr, _i64, _i64, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr, _i64, _i32,
                                         _ptr, _i64, _ptr, _i32, _i32, _ptr, _i64, _i32, _ptr, _i32, _ptr, _ptr]),
    "hgin_linear_fwd_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "hgin_linear_fwd": (_i32, [_i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _ptr, _i32, _i32, _ptr, _ptr, _i64,
                               _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr]),
    "hgin_linear_bwd_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "hgin_linear_bwd": (_i32, [_i64, _ptr, _i64, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr,
                               _i32, _i32, _i32, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32,
                               _ptr]),
    "hgin_linear_bwd_post": (_i32, [_i64, _ptr, _i64, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr,
                                    _i32, _i32, _i32, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr,
                                    _ptr, _i64, _i32, _ptr]),
    "hgin_linear_bwd_post_self": (_i32, [_i64, _ptr, _i64, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _i32, _ptr, _i64,
                                         _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32,
                                         _ptr]),
    "hgin_linear_fwd_t": (_i32, [_i32, _i32, _i64, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _ptr, _i32, _i32, _ptr, _ptr, _i64,
                                 _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr]),
    "hgin_linear_bwd_t": (_i32, [_i32, _i32, _i64, _ptr, _i64, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _i64, _i32, _ptr,
                                 _i32, _i32, _i32, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr,
                                 _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr]),
    "hgin_debug_gemm_tn_bf16": (_i32, [_i64, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr]),
    "hgin_reduce_workspace_bytes": (_i64, [_i64]),
    "hgin_mape_sum": (_i32, [_i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_sqrt_mape_bwd": (_i32, [_i64, _ptr, _ptr, _ptr, _f32, _ptr, _ptr, _ptr]),
    "hgin_adam_step": (_i32, [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _f64, _f64, _f64, _f64, _f64, _i32, _ptr]),
    "hgin_increment": (_i32, [_ptr, _ptr]),
    "hgin_collate_offsets": (_i32, [_i32, _ptr, _i32, _ptr, _i64, _ptr, _ptr, _ptr]),
    "hgin_collate_gather": (_i32, [_i32, _ptr, _i64, _i32, ctypes.POINTER(CollateField), _i32, _ptr, _i64, _ptr]),
    "hgin_qt_baseline_workspace_bytes": (_i64, [_i64, _i64]),
    "hgin_qt_baseline": (_i32, [_i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr,
                                _i64, _ptr]),
    "hgin_host_collate": (_i32, [_i32, _ptr, _i64, _i32, ctypes.POINTER(CollateField), _i32, _ptr, _ptr, _i32]),
    "hgin_elementwise_workspace_bytes": (_i64, []),
    "hgin_act_fwd": (_i32, [_i32, _i64, _i32, _ptr, _i64, _i32, _ptr, _f32, _f32, _ptr, _i64, _ptr]),
    "hgin_act_bwd": (_i32, [_i32, _i64, _i32, _ptr, _i64, _ptr, _i64, _i32, _ptr, _f32, _f32, _ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    "hgin_dropout": (_i32, [_i32, _i64, _i32, _ptr, _i64, _f32, _u64, _u64, _ptr, _i64, _ptr]),
    "hgin_bn_workspace_bytes": (_i64, [_i64, _i32]),
    "hgin_bn_stats": (_i32, [_i32, _i64, _i32, _ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    "hgin_bn_finalize": (_i32, [_i32, _ptr, _f64, _f64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "hgin_bn_act_fwd": (_i32, [_i32, _i64, _i32, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _f32, _f32, _ptr, _i64, _ptr]),
    "hgin_bn_act_bwd_reduce": (_i32, [_i32, _i64, _i32, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _f32, _f32,
                                      _ptr, _ptr, _i64, _ptr]),
    "hgin_bn_act_bwd_apply": (_i32, [_i32, _i64, _i32, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _f32, _f32,
                                     _ptr, _f64, _i32, _ptr, _i64, _ptr, _ptr, _ptr, _ptr]),
    "hgin_segment_pool": (_i32, [_i64, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr]),
    "hgin_readout_tail": (_i32, [_i64, _ptr, _i32, _i64, _ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _i64, _ptr]),
    "hgin_gat_fwd": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _i32, _i32, _f32, _i32, _i32, _ptr, _i64, _ptr,
                            _ptr, _ptr]),
    "hgin_gat_bwd": (_i32, [_i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _f32,
                            _i32, _ptr, _i64, _ptr, _ptr, _ptr, _ptr]),
    "hgin_small_step_workspace_bytes": (_i64, [_i64]),
    "hgin_small_step": (_i32, [_i64, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _i32, _i32, _i32, _i32]
                        + [_ptr] * 11 + [_ptr] * 11 + [_ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_small_step_phase": (_i32, [_i32, _i64, _ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _i64, _i32, _ptr, _ptr, _i32, _i32, _i32, _i32]
                              + [_ptr] * 11 + [_ptr] * 11 + [_ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "hgin_set_option": (_i32, [ctypes.c_char_p, _i32]),
    "hgin_debug_gemm_tn": (_i32, [_i64, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _ptr]),
}

_lib = None


class HginError(RuntimeError):
    pass


def load():
    """Load libhgin.so once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HginError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C gnn_link_prediction_b200/csrc`). "
                "There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what=""):
    if status != OK:
        msg = load().hgin_last_error().decode(errors="replace")
        raise HginError(f"{what or 'libhgin'} failed with status {status}: {msg}")


def header_functions(path=HEADER_PATH):
    """Names of the functions include/hgin.h declares (used by the ABI test)."""
    with open(path) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(hgin_[a-z0-9_]+)\s*\(", text)))
