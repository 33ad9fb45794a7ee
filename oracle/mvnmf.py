"""
Oracle (test infrastructure): minimum-volume NMF numerics in numpy float64.

X (V, D), W (V, k), H (k, D) as in oracle/klnmf.py.  Written from SURVEY.md
Appendix A.3; each function names the reference lines it restates.
"""

from __future__ import annotations

import numpy as np

from . import EPSILON
from .klnmf import fit_loop, kl_divergence, update_H


def normalize_WH(W, H):
    """(W / colsum(W), H * colsum(W)[:, None]).  Restates reference utils.py:155-158."""
    s = W.sum(axis=0)
    return W / s, H * s[:, None]


def volume_logdet(W, delta) -> float:
    """ln det(W^T W + delta I) through an LU determinant (not Cholesky).

    Restates reference models/mvnmf.py:19-24.
    """
    k = W.shape[1]
    return float(np.log(np.linalg.det(W.T @ W + delta * np.eye(k))))


def kl_divergence_penalized(X, W, H, lam, delta) -> float:
    """KL + lam * logdet volume.  Restates reference models/mvnmf.py:27-34."""
    return kl_divergence(X, W, H) + lam * volume_logdet(W, delta)


def update_W_unconstrained(X, W, H, lam, delta, n_given_signatures=0) -> np.ndarray:
    """Closed-form positive root of the volume-majorised W problem.

    Restates reference models/mvnmf.py:37-66.
    """
    X, W, H = (np.asarray(a, dtype=np.float64) for a in (X, W, H))
    k = W.shape[1]
    Y = np.linalg.inv(W.T @ W + delta * np.eye(k))
    Ym = np.maximum(0.0, -Y)
    Ya = np.abs(Y)
    WYm = W @ Ym
    WYa = W @ Ya
    r = H.sum(axis=1)
    N = (X / (W @ H)) @ H.T
    s1 = (r - 4.0 * lam * WYm) ** 2
    s2 = 8.0 * lam * WYa * N
    num = np.sqrt(s1 + s2) + (-r + 4.0 * lam * WYm)
    Wu = W * num / (4.0 * lam * WYa)
    g = n_given_signatures
    Wu[:, :g] = W[:, :g]
    Wu[:, g:] = np.clip(Wu[:, g:], EPSILON, None)
    return Wu


def line_search(X, W, H, lam, delta, gamma, W_unconstrained):
    """Back-tracking on the penalised objective; first trial ignores gamma.

    Restates reference models/mvnmf.py:69-92.  Returns (W_new, H_new, gamma).
    """
    prev = kl_divergence_penalized(X, W, H, lam, delta)
    Wn, Hn = normalize_WH(W_unconstrained, H)
    Wn, Hn = np.clip(Wn, EPSILON, None), np.clip(Hn, EPSILON, None)
    val = kl_divergence_penalized(X, Wn, Hn, lam, delta)
    while val > prev and gamma > 1e-16:
        gamma *= 0.8
        Wn = (1.0 - gamma) * W + gamma * W_unconstrained
        Wn, Hn = normalize_WH(Wn, H)
        Wn, Hn = np.clip(Wn, EPSILON, None), np.clip(Hn, EPSILON, None)
        val = kl_divergence_penalized(X, Wn, Hn, lam, delta)
    gamma = min(1.0, 1.2 * gamma)
    return Wn, Hn, gamma


def mvnmf_iteration(X, W, H, lam, delta, gamma, n_given_signatures=0):
    """One MvNMF._update_parameters.  Restates reference models/mvnmf.py:190-210."""
    H = update_H(X, W, H)
    if n_given_signatures == W.shape[1]:
        return W, H, gamma
    Wu = update_W_unconstrained(X, W, H, lam, delta, n_given_signatures)
    return line_search(X, W, H, lam, delta, gamma, Wu)


def fit_mvnmf(
    X,
    W0,
    H0,
    lam=1.0,
    delta=1.0,
    n_given_signatures=0,
    min_iterations=500,
    max_iterations=10000,
    conv_test_freq=10,
    tol=1e-7,
):
    """MvNMF.fit from a given start (gamma reset to 1, reference mvnmf.py:212-218).

    Returns (W, H, gamma, n_iterations, history).
    """
    X = np.asarray(X, dtype=np.float64)
    st = {"W": np.array(W0, dtype=np.float64), "H": np.array(H0, dtype=np.float64), "g": 1.0}

    def step():
        st["W"], st["H"], st["g"] = mvnmf_iteration(X, st["W"], st["H"], lam, delta, st["g"], n_given_signatures)

    def objective():
        return kl_divergence_penalized(X, st["W"], st["H"], lam, delta)

    n, hist = fit_loop(step, objective, min_iterations, max_iterations, conv_test_freq, tol)
    return st["W"], st["H"], st["g"], n, hist
