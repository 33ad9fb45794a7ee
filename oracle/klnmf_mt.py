"""
Oracle (test infrastructure): the reference's joint KL-NMF step and KL objective, restated so
that ALL host cores are used -- the CPU arm of bench.py (``--impl reference`` and the
``cpu_baseline`` leg).  Same arithmetic as ``oracle.klnmf.update_WH`` / ``kl_divergence``
(reference models/_utils_klnmf.py:281-361 and :11-55), float64, but the sample axis is cut
into chunks that a thread pool works through (numpy releases the GIL inside BLAS and ufuncs).
The partial W numerators are summed in chunk order, so the result equals the single-threaded
oracle up to the summation order over samples; tests/test_oracle_mt.py pins that.

Arrays use the AnnData memory layout of the reference: X [D][V], H [D][k], W [k][V].
"""

from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import EPSILON

try:  # keep BLAS single-threaded inside the workers; the pool supplies the parallelism
    from threadpoolctl import threadpool_limits
except Exception:  # pragma: no cover
    threadpool_limits = None


def n_host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def _chunks(D: int, n_threads: int, chunk: int | None):
    if chunk is None:
        chunk = max(1024, min(16384, -(-D // (4 * n_threads))))
    return [(lo, min(D, lo + chunk)) for lo in range(0, D, chunk)]


class HostKLNMF:
    """Joint multiplicative updates on the host with a persistent thread pool."""

    def __init__(self, X: np.ndarray, n_threads: int | None = None, chunk: int | None = None):
        self.X = np.ascontiguousarray(X, dtype=np.float64)  # [D][V]
        self.n_threads = n_threads or n_host_threads()
        self.pool = ThreadPoolExecutor(self.n_threads)
        self.bounds = _chunks(self.X.shape[0], self.n_threads, chunk)

    def close(self):
        self.pool.shutdown()

    def _map(self, fn):
        if threadpool_limits is not None:
            with threadpool_limits(limits=1):
                return list(self.pool.map(fn, self.bounds))
        return list(self.pool.map(fn, self.bounds))

    def update_WH(self, W: np.ndarray, H: np.ndarray, n_given_signatures: int = 0):
        """One joint step; W [k][V] is returned new, H [D][k] is updated IN PLACE (old W, reference :345)."""
        X = self.X
        k = W.shape[0]

        def work(b):
            lo, hi = b
            A = X[lo:hi] / (H[lo:hi] @ W)  # [d][V]
            num = H[lo:hi].T @ A  # [k][V]
            np.multiply(H[lo:hi], A @ W.T, out=H[lo:hi])
            np.maximum(H[lo:hi], EPSILON, out=H[lo:hi])
            return num

        parts = self._map(work)
        if n_given_signatures == k:
            return W, H
        num = parts[0].copy()
        for p in parts[1:]:
            num += p
        Wn = W * num
        Wn /= Wn.sum(axis=1, keepdims=True)
        Wn[:n_given_signatures] = W[:n_given_signatures]
        return np.maximum(Wn, EPSILON), H

    def kl_divergence(self, W: np.ndarray, H: np.ndarray) -> float:
        X = self.X

        def work(b):
            lo, hi = b
            x = X[lo:hi]
            WH = H[lo:hi] @ W
            nz = x != 0
            safe = np.where(nz, x, 1.0)
            return float((np.where(nz, x * np.log(safe / WH) - x, 0.0) + WH).sum())

        return float(sum(self._map(work)))
