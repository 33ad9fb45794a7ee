"""
Thin Python wrapper over the C ABI: one ``Workspace`` per (V, D_local, k, dtype) problem.

PyTorch owns device memory and streams (plumbing); every numerical step is a call into
libsalamander_b200.so through ctypes with raw device pointers.  No fallback exists: a
missing library or a CPU tensor raises.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (  # noqa: F401  (re-exported for callers)
    PASS_HSUM,
    PASS_NOCLIP,
    PASS_OBJECTIVE,
    PASS_POISSON,
    PASS_SAMPLEWISE,
    PASS_SCALED_UPDATE,
    PASS_UPDATE_H,
    PASS_WNUM,
)

_DTYPES = {torch.float32: _lib.SAL_F32, torch.float64: _lib.SAL_F64}


def resolve_dtype(dtype) -> torch.dtype:
    if isinstance(dtype, torch.dtype):
        out = dtype
    else:
        out = {"float64": torch.float64, "fp64": torch.float64, "float32": torch.float32, "fp32": torch.float32}.get(
            str(dtype)
        )
    if out not in _DTYPES:
        raise ValueError("dtype has to be 'float64' or 'float32'.")
    return out


def resolve_device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.SalamanderB200Error(
            "salamander_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise _lib.SalamanderB200Error(f"salamander_b200 runs on CUDA devices only, got '{dev}'.")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class Workspace:
    """Owns a ``sal_handle_t``; methods map 1:1 onto the ABI entry points."""

    def __init__(self, V: int, D_local: int, k: int, dtype: torch.dtype, device: torch.device, math: str = "fma"):
        self.lib = _lib.load()
        self.V, self.D, self.k = int(V), int(D_local), int(k)
        self.dtype, self.device = dtype, device
        self._h = C.c_void_p()
        self.timing = False
        _lib.check(
            self.lib.sal_create(C.byref(self._h), self.V, self.D, self.k, _DTYPES[dtype], device.index), "sal_create"
        )
        self.set_math(math)

    def set_math(self, math: str) -> None:
        mode = {"fma": _lib.MATH_FMA, "tf32": _lib.MATH_TF32, "tf32_always": _lib.MATH_TF32_ALWAYS}.get(math)
        if mode is None:
            raise ValueError("math has to be 'fma', 'tf32' or 'tf32_always'.")
        _lib.check(self.lib.sal_set_math(self._h, mode), "sal_set_math")
        self.math = math

    def set_timing(self, on: bool) -> None:
        _lib.check(self.lib.sal_set_timing(self._h, int(bool(on))), "sal_set_timing")
        self.timing = bool(on)

    def pass_timing(self) -> tuple[float, int]:
        """(summed milliseconds, count) of the UPDATE_H | WNUM pass kernels since the last call (CUDA events)."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        _lib.check(self.lib.sal_get_pass_timing(self._h, C.byref(ms), C.byref(n)), "sal_get_pass_timing")
        return float(ms.value), int(n.value)

    def set_debug_buffer(self, buf) -> None:
        """Diagnostics of the tensor-core pass (sal_set_debug_buffer); ``None`` switches them off."""
        ptr = None if buf is None else C.c_void_p(buf.data_ptr())
        _lib.check(self.lib.sal_set_debug_buffer(self._h, ptr), "sal_set_debug_buffer")
        self._dbg = buf  # keep the tensor alive while the library holds its address

    def close(self) -> None:
        if self._h:
            self.lib.sal_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.sal_launch_count(self._h))

    # -- helpers -----------------------------------------------------------------------------
    def _ptr(self, t: Optional[torch.Tensor], numel: int, name: str, dtype=None):
        if t is None:
            return None
        want = self.dtype if dtype is None else dtype
        if not isinstance(t, torch.Tensor) or t.device != self.device:
            raise ValueError(f"'{name}' has to be a tensor on {self.device}.")
        if t.dtype != want or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"'{name}' has to be a contiguous {want} tensor with {numel} elements.")
        return C.c_void_p(t.data_ptr())

    def _objectives_ptr(self, t: torch.Tensor):
        """The objectives of the period kernel may be written straight into PAGE-LOCKED host memory (unified addressing: the
        kernel's few 8-byte stores travel over PCIe and are visible once the launch has completed), which saves the
        device-to-host copy behind every launch; a device tensor is accepted as well."""
        if isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.is_pinned() and t.dtype == torch.float64 and t.is_contiguous():
            return C.c_void_p(t.data_ptr())
        return self._ptr(t, t.numel(), "objectives", torch.float64)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- ABI ---------------------------------------------------------------------------------
    def klnmf_pass(
        self,
        X,
        W,
        H_in,
        flags: int,
        H_out=None,
        w_kl=None,
        w_lhalf=None,
        h_scale=None,
        Wnum=None,
        objective=None,
        per_sample=None,
        hsum=None,
    ) -> None:
        V, D, k = self.V, self.D, self.k
        _lib.check(
            self.lib.sal_klnmf_pass(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W, k * V, "W"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                self._ptr(w_kl, D, "w_kl"),
                self._ptr(w_lhalf, D, "w_lhalf"),
                self._ptr(h_scale, k, "h_scale"),
                int(flags),
                self._ptr(Wnum, k * V, "Wnum"),
                self._ptr(objective, 1, "objective", torch.float64),
                self._ptr(per_sample, D, "per_sample"),
                self._ptr(hsum, k, "hsum"),
                self._stream(),
            ),
            "sal_klnmf_pass",
        )

    def klnmf_update(self, X, W_in, W_out, H_in, H_out, n_given: int, clip_given: bool, Wnum, w_kl=None, w_lhalf=None, objective=None) -> None:
        """Joint update in two launches: fused pass, then reduction + W epilogue (sal_klnmf_update)."""
        V, D, k = self.V, self.D, self.k
        _lib.check(
            self.lib.sal_klnmf_update(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W_in, k * V, "W_in"),
                self._ptr(W_out, k * V, "W_out"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                self._ptr(w_kl, D, "w_kl"),
                self._ptr(w_lhalf, D, "w_lhalf"),
                int(n_given),
                int(bool(clip_given)),
                self._ptr(Wnum, k * V, "Wnum"),
                self._ptr(objective, 1, "objective", torch.float64),
                self._stream(),
            ),
            "sal_klnmf_update",
        )

    def klnmf_update_p2p(self, X, W_in, W_out, H_in, H_out, n_given: int, clip_given: bool, Wnum, peers, state, n_ranks: int, rank: int,
                         w_kl=None, w_lhalf=None, objective=None) -> None:
        """Multi-GPU joint update: fused pass, then reduction + one-shot NVLink all-reduce + W epilogue in one kernel."""
        V, D, k = self.V, self.D, self.k
        _lib.check(
            self.lib.sal_klnmf_update_p2p(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W_in, k * V, "W_in"),
                self._ptr(W_out, k * V, "W_out"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                self._ptr(w_kl, D, "w_kl"),
                self._ptr(w_lhalf, D, "w_lhalf"),
                int(n_given),
                int(bool(clip_given)),
                self._ptr(Wnum, k * V, "Wnum"),
                self._ptr(objective, 1, "objective", torch.float64),
                self._ptr(peers, n_ranks, "peers", torch.int64),
                self._ptr(state, 2, "state", torch.int32),
                int(n_ranks),
                int(rank),
                self._stream(),
            ),
            "sal_klnmf_update_p2p",
        )

    def period_supported(self, n_given: int = 0, n_ranks: int = 1) -> bool:
        return bool(self.lib.sal_klnmf_period_supported(self._h, int(n_given), int(n_ranks)))

    def klnmf_period(self, X, W_in, W_out, H_in, H_out, n_given: int, clip_given: bool, n_updates: int, objective_every: int,
                     final_objective: bool, objectives=None, peers=None, state=None, n_ranks: int = 1, rank: int = 0) -> None:
        """``n_updates`` joint updates (+ fused / trailing objectives) in ONE persistent launch (sal_klnmf_period)."""
        V, D, k = self.V, self.D, self.k
        n_obj = (-(-int(n_updates) // int(objective_every)) if objective_every else 0) + int(bool(final_objective))
        if n_obj and (objectives is None or objectives.numel() < n_obj):
            raise ValueError(f"'objectives' has to hold {n_obj} doubles.")
        _lib.check(
            self.lib.sal_klnmf_period(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W_in, k * V, "W_in"),
                self._ptr(W_out, k * V, "W_out"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                int(n_given),
                int(bool(clip_given)),
                int(n_updates),
                int(objective_every),
                int(bool(final_objective)),
                None if objectives is None else self._objectives_ptr(objectives),
                None if peers is None else self._ptr(peers, n_ranks, "peers", torch.int64),
                None if state is None else self._ptr(state, 2, "state", torch.int32),
                int(n_ranks),
                int(rank),
                self._stream(),
            ),
            "sal_klnmf_period",
        )

    def small_supported(self) -> bool:
        return bool(self.lib.sal_klnmf_small_supported(self._h))

    def klnmf_small_updates(self, X, W_in, W_out, H_in, H_out, n_given: int, n_iterations: int, objective=None) -> None:
        """``n_iterations`` joint updates in one launch of a single persistent CTA (sal_klnmf_small_updates)."""
        V, D, k = self.V, self.D, self.k
        _lib.check(
            self.lib.sal_klnmf_small_updates(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W_in, k * V, "W_in"),
                self._ptr(W_out, k * V, "W_out"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                int(n_given),
                int(n_iterations),
                self._ptr(objective, 1, "objective", torch.float64),
                self._stream(),
            ),
            "sal_klnmf_small_updates",
        )

    def mvnmf_small_supported(self) -> bool:
        return bool(self.lib.sal_mvnmf_small_supported(self._h))

    def mvnmf_small_updates(self, X, W_in, W_out, H_in, H_out, lam: float, delta: float, n_given: int, n_iterations: int,
                            gamma_in, gamma_out, objective=None) -> None:
        """``n_iterations`` whole MvNMF iterations (line search included) in one launch of a single persistent CTA."""
        V, D, k = self.V, self.D, self.k
        _lib.check(
            self.lib.sal_mvnmf_small_updates(
                self._h,
                self._ptr(X, D * V, "X"),
                self._ptr(W_in, k * V, "W_in"),
                self._ptr(W_out, k * V, "W_out"),
                self._ptr(H_in, D * k, "H_in"),
                self._ptr(H_out, D * k, "H_out"),
                float(lam),
                float(delta),
                int(n_given),
                int(n_iterations),
                self._ptr(gamma_in, 1, "gamma_in", torch.float64),
                self._ptr(gamma_out, 1, "gamma_out", torch.float64),
                self._ptr(objective, 1, "objective", torch.float64),
                self._stream(),
            ),
            "sal_mvnmf_small_updates",
        )

    def w_epilogue(self, W_in, Wnum, n_given: int, clip_given: bool, W_out) -> None:
        kv = self.k * self.V
        _lib.check(
            self.lib.sal_w_epilogue(
                self._h,
                self._ptr(W_in, kv, "W_in"),
                self._ptr(Wnum, kv, "Wnum"),
                int(n_given),
                int(bool(clip_given)),
                self._ptr(W_out, kv, "W_out"),
                self._stream(),
            ),
            "sal_w_epilogue",
        )

    def clip_counts(self, X, n_changed) -> None:
        """X <- max(X, EPSILON) in place; ``n_changed`` (zeroed int64 device scalar) counts raised entries."""
        _lib.check(
            self.lib.sal_clip_counts(
                self._h,
                self._ptr(X, X.numel(), "X"),
                int(X.numel()),
                self._ptr(n_changed, 1, "n_changed", torch.int64),
                self._stream(),
            ),
            "sal_clip_counts",
        )

    def scale_clip_rows(self, H, scale) -> None:
        """H[d][j] <- max(H[d][j] * scale[j], EPSILON) in place (sal_scale_clip_rows)."""
        _lib.check(
            self.lib.sal_scale_clip_rows(self._h, self._ptr(H, self.D * self.k, "H"), self._ptr(scale, self.k, "scale"), self._stream()),
            "sal_scale_clip_rows",
        )

    def mvnmf_logdet(self, W, delta: float, out) -> None:
        _lib.check(
            self.lib.sal_mvnmf_logdet(
                self._h, self._ptr(W, self.k * self.V, "W"), float(delta), self._ptr(out, 1, "out", torch.float64), self._stream()
            ),
            "sal_mvnmf_logdet",
        )

    def mvnmf_w_unconstrained(self, W, N, hsum, lam: float, delta: float, n_given: int, W_unc) -> None:
        kv = self.k * self.V
        _lib.check(
            self.lib.sal_mvnmf_w_unconstrained(
                self._h,
                self._ptr(W, kv, "W"),
                self._ptr(N, kv, "N"),
                self._ptr(hsum, self.k, "hsum"),
                float(lam),
                float(delta),
                int(n_given),
                self._ptr(W_unc, kv, "W_unc"),
                self._stream(),
            ),
            "sal_mvnmf_w_unconstrained",
        )

    def mvnmf_w_unconstrained_trial(self, W, N, hsum, lam: float, delta: float, n_given: int, W_unc, W_trial, h_scale, logdet_out) -> None:
        """W_unconstrained and the first line-search candidate (the full step) in one launch."""
        kv = self.k * self.V
        _lib.check(
            self.lib.sal_mvnmf_w_unconstrained_trial(
                self._h, self._ptr(W, kv, "W"), self._ptr(N, kv, "N"), self._ptr(hsum, self.k, "hsum"), float(lam), float(delta),
                int(n_given), self._ptr(W_unc, kv, "W_unc"), self._ptr(W_trial, kv, "W_trial"), self._ptr(h_scale, self.k, "h_scale"),
                self._ptr(logdet_out, 1, "logdet_out", torch.float64), self._stream(),
            ),
            "sal_mvnmf_w_unconstrained_trial",
        )

    def mvnmf_trial(self, W, W_unc, gamma_blend: float, delta: float, W_trial, h_scale, logdet_out) -> None:
        kv = self.k * self.V
        _lib.check(
            self.lib.sal_mvnmf_trial(
                self._h,
                self._ptr(W, kv, "W"),
                self._ptr(W_unc, kv, "W_unc"),
                float(gamma_blend),
                float(delta),
                self._ptr(W_trial, kv, "W_trial"),
                self._ptr(h_scale, self.k, "h_scale"),
                self._ptr(logdet_out, 1, "logdet_out", torch.float64),
                self._stream(),
            ),
            "sal_mvnmf_trial",
        )


def klnmf_period_emulated(workspaces, Xs, W_ins, W_outs, H_ins, H_outs, n_given: int, clip_given: bool, n_updates: int,
                          objective_every: int, final_objective: bool, objectives, peer_tables, states) -> None:
    """1 or 2 emulated ranks (one Workspace, shard and receive buffer each) inside ONE cooperative launch on one GPU
    (sal_klnmf_period_emulated): the multi-GPU exchange protocol of the period kernel without a second GPU."""
    n = len(workspaces)
    lib = workspaces[0].lib

    def arr(ptrs):
        return (C.c_void_p * n)(*[p.value if isinstance(p, C.c_void_p) else p for p in ptrs])

    hs = arr([w._h for w in workspaces])
    cols = []
    for name, ts, per_el in (("X", Xs, "DV"), ("W_in", W_ins, "kV"), ("W_out", W_outs, "kV"), ("H_in", H_ins, "Dk"), ("H_out", H_outs, "Dk")):
        ptrs = []
        for w, t in zip(workspaces, ts):
            numel = {"DV": w.D * w.V, "kV": w.k * w.V, "Dk": w.D * w.k}[per_el]
            ptrs.append(w._ptr(t, numel, name))
        cols.append(arr(ptrs))
    objs = arr([w._ptr(o, o.numel(), "objectives", torch.float64) for w, o in zip(workspaces, objectives)])
    tabs = arr([w._ptr(t, n, "peer_table", torch.int64) for w, t in zip(workspaces, peer_tables)])
    sts = arr([w._ptr(t, 2, "state", torch.int32) for w, t in zip(workspaces, states)])
    _lib.check(
        lib.sal_klnmf_period_emulated(
            hs, n, cols[0], cols[1], cols[2], cols[3], cols[4], int(n_given), int(bool(clip_given)), int(n_updates),
            int(objective_every), int(bool(final_objective)), objs, tabs, sts, workspaces[0]._stream(),
        ),
        "sal_klnmf_period_emulated",
    )


def corrnmf_signature_embeddings_emulated(workspaces, auxTs, a_s, bs, Ls, Us, m: int, variance: float, peer_tables, launch_id: int) -> None:
    """1 or 2 emulated ranks (one Workspace, sample shard and receive buffer each) inside ONE launch on one GPU
    (sal_corrnmf_signature_embeddings_emulated): the in-kernel exchange of the signature-embedding solver without a second GPU."""
    n = len(workspaces)
    lib = workspaces[0].lib

    def arr(tensors):
        return (C.c_void_p * n)(*[int(t.data_ptr()) for t in tensors])

    hs = (C.c_void_p * n)(*[w._h.value if isinstance(w._h, C.c_void_p) else w._h for w in workspaces])
    _lib.check(
        lib.sal_corrnmf_signature_embeddings_emulated(
            hs, n, arr(auxTs), arr(a_s), arr(bs), arr(Ls), arr(Us), int(m), float(variance), arr(peer_tables), int(launch_id),
            workspaces[0]._stream(),
        ),
        "sal_corrnmf_signature_embeddings_emulated",
    )


def trim_device_scratch() -> None:
    """Free the workspace scratch buffers that finished fits left on the library's per-device free list (sal_trim_scratch)."""
    _lib.check(_lib.load().sal_trim_scratch(), "sal_trim_scratch")
