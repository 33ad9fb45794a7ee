"""
Oracle (test infrastructure): KL-NMF numerics in numpy float64.

Conventions follow the reference's free functions: X is (V, D) features x samples,
W is (V, k) with columns ~ summing to one, H is (k, D).  Written from SURVEY.md
Appendix A.1/A.2; each function names the reference lines it restates.
"""

from __future__ import annotations

import numpy as np
from scipy.special import gammaln

from . import EPSILON


def _as_f64(*arrays):
    return tuple(np.asarray(a, dtype=np.float64) for a in arrays)


def kl_divergence(X, W, H, weights=None) -> float:
    """Generalised KL divergence D(X || WH), optionally sample-weighted.

    Restates reference models/_utils_klnmf.py:11-55: cells with X == 0 contribute
    only (WH); every other cell contributes x ln(x/wh) - x + wh.
    """
    X, W, H = _as_f64(X, W, H)
    WH = W @ H
    nz = X != 0
    safe_x = np.where(nz, X, 1.0)
    cell = np.where(nz, X * np.log(safe_x / WH) - X, 0.0) + WH
    per_sample = cell.sum(axis=0)
    if weights is not None:
        per_sample = per_sample * np.asarray(weights, dtype=np.float64)
    return float(per_sample.sum())


def samplewise_kl_divergence(X, W, H, weights=None) -> np.ndarray:
    """Per-sample KL with zero cells replaced by EPSILON in both X and WH.

    Restates reference models/_utils_klnmf.py:58-97 (s1 + s2 + s3 with
    s3 = H^T colsum(W)).
    """
    X, W, H = _as_f64(X, W, H)
    zero = X == 0
    Xe = np.where(zero, EPSILON, X)
    WHe = np.where(zero, EPSILON, W @ H)
    s1 = (Xe * np.log(Xe / WHe)).sum(axis=0)
    s2 = -X.sum(axis=0)
    s3 = H.T @ W.sum(axis=0)
    out = s1 + s2 + s3
    if weights is not None:
        out = out * np.asarray(weights, dtype=np.float64)
    return out


def poisson_llh(X, W, H) -> float:
    """Generalised Poisson log-likelihood sum x ln(wh) - wh - lnGamma(1 + x).

    Restates reference models/_utils_klnmf.py:100-161.
    """
    X, W, H = _as_f64(X, W, H)
    WH = W @ H
    nz = WH != 0
    val = np.where(nz, X * np.log(np.where(nz, WH, 1.0)), 0.0) - WH
    return float(val.sum() - gammaln(1.0 + X).sum())


def _ratio(X, W, H):
    return X / (W @ H)


def _h_step(H, WtA, weights_kl, weights_lhalf):
    """H half shared by update_H / update_WH (reference :258-278, :343-361)."""
    if weights_lhalf is None:
        return np.clip(H * WtA, EPSILON, None)
    lam = np.asarray(weights_lhalf, dtype=np.float64)
    t = 4.0 * H * WtA
    if weights_kl is not None:
        wsq = np.asarray(weights_kl, dtype=np.float64) ** 2
        t = t * wsq
    disc = 0.25 * lam**2 + t
    Hn = 0.25 * (lam / 2.0 - np.sqrt(disc)) ** 2
    if weights_kl is not None:
        Hn = Hn / wsq
    return np.clip(Hn, EPSILON, None)


def update_W(X, W, H, weights_kl=None, n_given_signatures=0) -> np.ndarray:
    """W multiplicative step; clips only the NON-given columns.

    Restates reference models/_utils_klnmf.py:164-217.
    """
    X, W, H = _as_f64(X, W, H)
    k = W.shape[1]
    if n_given_signatures == k:
        return W
    A = _ratio(X, W, H)
    if weights_kl is not None:
        A = A * np.asarray(weights_kl, dtype=np.float64)
    Wn = W * (A @ H.T)
    Wn = Wn / Wn.sum(axis=0)
    g = n_given_signatures
    Wn[:, :g] = W[:, :g]
    Wn[:, g:] = np.clip(Wn[:, g:], EPSILON, None)
    return Wn


def update_H(X, W, H, weights_kl=None, weights_lhalf=None) -> np.ndarray:
    """H multiplicative step (or l-half closed form).

    Restates reference models/_utils_klnmf.py:220-278.  Returns a new array
    (the reference mutates H in place in the un-penalised branch).
    """
    X, W, H = _as_f64(X, W, H)
    A = _ratio(X, W, H)
    return _h_step(H, W.T @ A, weights_kl, weights_lhalf)


def update_WH(X, W, H, weights_kl=None, weights_lhalf=None, n_given_signatures=0):
    """Joint step sharing A = X/(WH); H uses the OLD W; ALL W columns are clipped.

    Restates reference models/_utils_klnmf.py:281-361.
    """
    X, W, H = _as_f64(X, W, H)
    k = W.shape[1]
    A = _ratio(X, W, H)
    if n_given_signatures == k:
        Wn = W
    else:
        As = A if weights_kl is None else A * np.asarray(weights_kl, dtype=np.float64)
        Wn = W * (As @ H.T)
        Wn = Wn / Wn.sum(axis=0)
        Wn[:, :n_given_signatures] = W[:, :n_given_signatures]
        Wn = np.clip(Wn, EPSILON, None)
    Hn = _h_step(H, W.T @ A, weights_kl, weights_lhalf)
    return Wn, Hn


def klnmf_objective(X, W, H, weights_kl=None, weights_lhalf=None) -> float:
    """KLNMF.objective_function: KL (+ sum_d lam_d sum_k sqrt(H_kd)).

    Restates reference models/klnmf.py:64-80.
    """
    val = kl_divergence(X, W, H, weights_kl)
    if weights_lhalf is not None:
        val += float(
            np.dot(np.asarray(weights_lhalf, dtype=np.float64), np.sqrt(np.asarray(H, dtype=np.float64)).sum(axis=0))
        )
    return val


def fit_loop(step, objective, min_iterations=500, max_iterations=10000, conv_test_freq=10, tol=1e-7):
    """The reference's convergence loop, restated from models/signature_nmf.py:358-385.

    ``step()`` performs one parameter update, ``objective()`` evaluates the current
    objective.  Returns (n_iterations, history) where history excludes the value
    computed before the first update (the reference drops ``of_values[0]``).
    """
    of_values = [objective()]
    n = 0
    converged = False
    while not converged:
        n += 1
        step()
        if n % conv_test_freq == 0:
            prev = of_values[-1]
            of_values.append(objective())
            rel = abs(prev - of_values[-1]) / abs(prev)
            converged = rel < tol and n >= min_iterations
        converged = converged or n >= max_iterations
    return n, of_values[1:]


def fit_klnmf(
    X,
    W0,
    H0,
    weights_kl=None,
    weights_lhalf=None,
    n_given_signatures=0,
    min_iterations=500,
    max_iterations=10000,
    conv_test_freq=10,
    tol=1e-7,
):
    """KLNMF.fit from a given (already normalised/clipped) start.

    X must already be clipped to EPSILON (reference models/signature_nmf.py:281).
    Returns (W, H, n_iterations, history).
    """
    state = {"W": np.array(W0, dtype=np.float64), "H": np.array(H0, dtype=np.float64)}
    X = np.asarray(X, dtype=np.float64)

    def step():
        state["W"], state["H"] = update_WH(X, state["W"], state["H"], weights_kl, weights_lhalf, n_given_signatures)

    def objective():
        return klnmf_objective(X, state["W"], state["H"], weights_kl, weights_lhalf)

    n, hist = fit_loop(step, objective, min_iterations, max_iterations, conv_test_freq, tol)
    return state["W"], state["H"], n, hist
