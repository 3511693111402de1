"""ctypes binding of librvcb200.so (the C ABI declared in include/rvcb200.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load the
import raises, and every wrapper raises ``RuntimeError`` with ``rvcb_last_error()`` on a
non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librvcb200.so")


class GemmSeg(C.Structure):
    _fields_ = [("row_off", C.c_int), ("col_off", C.c_int), ("dw", C.c_int), ("nk", C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_rows", C.c_int), ("a_cols", C.c_int), ("conv2d_W", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_rows", C.c_int), ("b_cols", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("block_k", C.c_int), ("nseg", C.c_int),
        ("batch", C.c_int), ("a_row_z", C.c_int64), ("a_col_z", C.c_int64), ("b_row_z", C.c_int64),
        ("b_col_z", C.c_int64), ("c_z", C.c_int64), ("bias_z", C.c_int64), ("b_col0", C.c_int),
        ("bias", C.c_void_p), ("bias_per_row", C.c_int),
        ("res1", C.c_void_p), ("ldres1", C.c_int64), ("res2", C.c_void_p), ("ldres2", C.c_int64),
        ("alpha", C.c_float), ("act1", C.c_int), ("act1_p", C.c_float), ("act2", C.c_int), ("act2_p", C.c_float),
        ("gate", C.c_int),
        ("out32", C.c_void_p), ("ld32", C.c_int64), ("out16", C.c_void_p), ("ld16", C.c_int64), ("up2_C", C.c_int),
        ("seg", GemmSeg * 128),
    ]


class SynthConfig(C.Structure):
    _fields_ = [
        ("inter_channels", C.c_int), ("hidden_channels", C.c_int), ("filter_channels", C.c_int),
        ("n_heads", C.c_int), ("n_layers", C.c_int), ("kernel_size", C.c_int),
        ("n_resblock_kernels", C.c_int), ("resblock_kernel_sizes", C.c_int * 4),
        ("resblock_dilations", (C.c_int * 3) * 4),
        ("n_upsamples", C.c_int), ("upsample_rates", C.c_int * 8), ("upsample_kernel_sizes", C.c_int * 8),
        ("upsample_initial_channel", C.c_int), ("spk_embed_dim", C.c_int), ("gin_channels", C.c_int),
        ("sr", C.c_int), ("encoder_dim", C.c_int),
    ]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol include/rvcb200.h declares
_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
SIGNATURES = {
    "rvcb_init": (_I, [_I]),
    "rvcb_last_error": (C.c_char_p, []),
    "rvcb_launch_count": (C.c_ulonglong, []),
    "rvcb_version": (C.c_char_p, []),
    "rvcb_set_grid_cap": (_I, [_I]),
    "rvcb_prof_begin": (_I, []),
    "rvcb_prof_end": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "rvcb_prof_classes": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rvcb_op_resblock1_out_rows": (_L, [_I, _I, C.POINTER(_I), _I]),
    "rvcb_op_resblock1": (_I, [_I, _I, C.POINTER(_I), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _I, _P, _P]),
    "rvcb_rt_tail": (_I, [_P, _I, _P, _I, _F, _P, _I, _I, _I, _P, _P, _P, _P]),
    "rvcb_rt_tail_pv": (_I, [_P, _I, _P, _I, _F, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rvcb_torchgate_create": (_I, [_I, _I, _I, _I, _F, _F, _F, _I, _F, _P, _I, _I, C.POINTER(_P)]),
    "rvcb_torchgate_out_len": (_L, [_P, _L]),
    "rvcb_torchgate_apply": (_I, [_P, _P, _L, _P, _L, _P, _P]),
    "rvcb_torchgate_destroy": (None, [_P]),
    "rvcb_resample_sinc": (_I, [_P, _L, _P, _I, _I, _I, _I, _P, _L, _P]),
    "rvcb_weights_create": (_I, [C.POINTER(_P)]),
    "rvcb_weights_add": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_L)]),
    "rvcb_weights_destroy": (None, [_P]),
    "rvcb_hubert_create": (_I, [_P, C.POINTER(_P)]),
    "rvcb_hubert_num_frames": (_I, [_I]),
    "rvcb_hubert_extract_features": (_I, [_P, _P, _I, _I, _P, C.POINTER(_I), _P]),
    "rvcb_hubert_final_proj": (_I, [_P, _P, _I, _P, _P]),
    "rvcb_hubert_destroy": (None, [_P]),
    "rvcb_index_create": (_I, [_P, _I, _P, _L, _I, _P, _P, C.POINTER(_P)]),
    "rvcb_index_ntotal": (_L, [_P]),
    "rvcb_index_search": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "rvcb_index_blend": (_I, [_P, _P, _I, _I, _P, _P, _F, _P, _P]),
    "rvcb_knn_bruteforce_top1": (_I, [_P, _L, _I, _P, _I, _P, _P, _P]),
    "rvcb_index_destroy": (None, [_P]),
    "rvcb_flat_create": (_I, [_P, _L, _I, C.POINTER(_P)]),
    "rvcb_flat_search_top1": (_I, [_P, _P, _I, _P, _P, _P]),
    "rvcb_flat_destroy": (None, [_P]),
    "rvcb_upsample_protect": (_I, [_P, _P, _I, _I, _P, _I, _F, _P, _P]),
    "rvcb_post_mix": (_I, [_P, _L, _I, _P, _L, _F, _P, _P]),
    "rvcb_rms_mix": (_I, [_P, _L, _I, _P, _L, _F, _P, _P]),
    "rvcb_host_filtfilt": (_I, [_P, _P, _P, _I, _P, _I, _L, _P]),
    "rvcb_sosfiltfilt": (_I, [_P, _P, _I, _I, _P, _L, _P, _P, _P]),
    "rvcb_reflect_pad": (_I, [_P, _L, _L, _P, _P]),
    "rvcb_f32_to_i16": (_I, [_P, _L, _P, _P]),
    "rvcb_f0_post": (_I, [_P, _I, _I, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P]),
    "rvcb_rmvpe_create": (_I, [_P, C.POINTER(_P)]),
    "rvcb_rmvpe_num_frames": (_I, [_I]),
    "rvcb_rmvpe_infer": (_I, [_P, _P, _I, _F, _P, _P, _P, C.POINTER(_I), _P]),
    "rvcb_rmvpe_destroy": (None, [_P]),
    "rvcb_synth_create": (_I, [C.POINTER(SynthConfig), _P, C.POINTER(_P)]),
    "rvcb_synth_infer": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _I, _P, C.POINTER(_I), _P]),
    "rvcb_synth_infer_keep": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P, C.POINTER(_I), _P]),
    "rvcb_synth_destroy": (None, [_P]),
    "rvcb_op_gemm": (_I, [C.POINTER(GemmDesc), _I, _P]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        l = C.CDLL(LIB_PATH)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(l, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing and not os.environ.get("RVCB_ALLOW_PARTIAL_LIB"):
            raise RuntimeError(f"{LIB_PATH} does not export: {missing} (stale build? run __graft_entry__.build())")
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise RuntimeError("librvcb200: " + lib().rvcb_last_error().decode("utf-8", "replace"))


_inited = {}


def init(device: int = 0) -> None:
    if device not in _inited:
        check(lib().rvcb_init(device))
        _inited[device] = True
