"""
ctypes binding of libsalamander_b200.so (the C ABI declared in include/salamander_b200.h).

There is NO CPU fallback: if the shared library is missing or the device is not a
Blackwell GPU every entry point raises.  PyTorch is used only to own device memory and
streams; the pointers handed to the library are ``tensor.data_ptr()``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsalamander_b200.so")

SAL_F32, SAL_F64 = 0, 1
MATH_FMA, MATH_TF32, MATH_TF32_ALWAYS = 0, 1, 2
PASS_UPDATE_H, PASS_WNUM, PASS_OBJECTIVE, PASS_SAMPLEWISE, PASS_HSUM, PASS_POISSON, PASS_NOCLIP = 1, 2, 4, 8, 16, 32, 64
PASS_PARTIALS_ONLY = 128
PASS_SCALED_UPDATE = 256

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double

#: every symbol include/salamander_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sal_last_error": (C.c_char_p, []),
    "sal_version": (_i, []),
    "sal_create": (_i, [C.POINTER(_vp), _i, _i64, _i, _i, _i]),
    "sal_destroy": (_i, [_vp]),
    "sal_trim_scratch": (_i, []),
    "sal_set_math": (_i, [_vp, _i]),
    "sal_launch_count": (_i64, [_vp]),
    "sal_set_timing": (_i, [_vp, _i]),
    "sal_get_pass_timing": (_i, [_vp, C.POINTER(_d), C.POINTER(_i64)]),
    "sal_set_debug_buffer": (_i, [_vp, _vp]),
    "sal_klnmf_pass": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sal_klnmf_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sal_klnmf_update_p2p": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "sal_p2p_exchange_bytes": (C.c_size_t, [_i, _i]),
    "sal_klnmf_period_supported": (_i, [_vp, _i, _i]),
    "sal_klnmf_period": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "sal_klnmf_period_emulated": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sal_klnmf_small_supported": (_i, [_vp]),
    "sal_klnmf_small_updates": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "sal_w_epilogue": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "sal_clip_counts": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "sal_mark_counts_written": (_i, [_vp]),
    "sal_scale_clip_rows": (_i, [_vp, _vp, _vp, _vp]),
    "sal_mvnmf_logdet": (_i, [_vp, _vp, _d, _vp, _vp]),
    "sal_mvnmf_w_unconstrained": (_i, [_vp, _vp, _vp, _vp, _d, _d, _i, _vp, _vp]),
    "sal_mvnmf_trial": (_i, [_vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp]),
    "sal_mvnmf_w_unconstrained_trial": (_i, [_vp, _vp, _vp, _vp, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "sal_mvnmf_small_supported": (_i, [_vp]),
    "sal_mvnmf_small_updates": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _vp, _vp, _vp, _vp]),
    "sal_corrnmf_max_dim": (_i, []),
    "sal_row_sums": (_i, [_vp, _vp, _vp, _vp]),
    "sal_corrnmf_exposures": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sal_corrnmf_sample_scalings": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sal_corrnmf_signature_scalings_sums": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sal_corrnmf_signature_scalings_finish": (_i, [_vp, _vp, _vp, _vp]),
    "sal_corrnmf_sample_embeddings": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _vp]),
    "sal_corrnmf_sample_embeddings_mm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _vp]),
    "sal_corrnmf_signature_embeddings": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _vp]),
    "sal_corrnmf_signature_embeddings_range": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _i, _vp]),
    "sal_corrnmf_sig_exchange_bytes": (C.c_size_t, [_i, _i]),
    "sal_p2p_allreduce_bytes": (C.c_size_t, [_i]),
    "sal_p2p_allreduce_max_values": (_i, []),
    "sal_p2p_allreduce_f64": (_i, [_vp, _vp, _i, _vp, _i, _i, C.c_uint, _vp]),
    "sal_p2p_allreduce_f64_emulated": (_i, [_vp, _vp, _i, _vp, _i, C.c_uint, _vp]),
    "sal_corrnmf_signature_embeddings_p2p": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _vp, _i, _i, C.c_uint, _vp]),
    "sal_corrnmf_signature_embeddings_emulated": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _d, _vp, C.c_uint, _vp]),
    "sal_corrnmf_norms": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
}

_lib = None


class SalamanderB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SalamanderB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  salamander_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    """Map the ABI's int status to Python exceptions (<0 ValueError/NotImplementedError, >0 CUDA)."""
    if status == 0:
        return
    msg = load().sal_last_error().decode(errors="replace")
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    if status < 0:
        raise ValueError(f"{what}: {msg}")
    raise SalamanderB200Error(f"{what}: CUDA error {status}: {msg}")
