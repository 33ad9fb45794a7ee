"""
Oracle (test infrastructure): correlated NMF numerics in numpy float64.

Restates reference models/_utils_corrnmf.py (exposures :11-25, aux :28-52, ELBO :55-100, scaling updates
:103-179, embedding objective / gradient / Hessian :182-351, update_embedding :354-410) and the update order of
models/corrnmf_det.py:157-169.  The embedding problems are minimised by
``scipy.optimize.minimize(method="Newton-CG")`` exactly as the reference does (:400-407; third-party dependency,
scipy 1.13.1 pinned by the reference's lock file, 1.18.1 in this image -- the reference's golden fixtures still
reproduce).  ``newton_cg`` below is a second, self-contained restatement of that published algorithm (truncated
Newton with a CG inner loop, More'-Thuente DCSRCH line search of MINPACK-2, SciPy's defaults); it is what the CUDA
kernels implement and tests/test_oracle_corrnmf.py pins it against scipy itself.

Conventions: data X (D, V); signatures W (k, V); exposures H (D, k); signature scalings a (k,), sample scalings
b (D,); signature embeddings L (k, m), sample embeddings U (D, m).
"""

from __future__ import annotations

import math

import numpy as np
from scipy import optimize
from scipy.special import gammaln

from . import EPSILON
from . import klnmf


def compute_exposures(a, b, L, U):
    """H_dk = exp(a_k + b_d + l_k . u_d)   (reference _utils_corrnmf.py:11-25)."""
    return np.exp(a[:, None] + b[None, :] + L @ U.T).T


def compute_aux(X, W, H):
    """aux_kd = H_dk * (W (X / (H W))^T)_kd   (reference :28-52)."""
    return H.T * (W @ (X / (H @ W)).T)


def poisson_llh(X, W, H):
    """sum x ln(wh) - wh - lnGamma(1 + x) in the (D, V) layout (reference _utils_klnmf.py:100-161)."""
    return klnmf.poisson_llh(X.T, W.T, H.T)


def elbo(X, W, H, L, U, variance, penalize_sample_embeddings=True):
    """Reference :55-100."""
    k, m = L.shape
    D = U.shape[0]
    val = poisson_llh(X, W, H)
    val -= 0.5 * m * k * np.log(2 * np.pi * variance)
    val -= np.sum(L**2) / (2 * variance)
    if penalize_sample_embeddings:
        val -= 0.5 * m * D * np.log(2 * np.pi * variance)
        val -= np.sum(U**2) / (2 * variance)
    return float(val)


def update_signature_scalings(aux, b, L, U):
    """a_k = ln sum_d aux_kd - ln sum_d exp(b_d + l_k . u_d)   (reference :103-138)."""
    return np.log(aux.sum(axis=1)) - np.log(np.exp(b[None, :] + L @ U.T).sum(axis=1))


def update_sample_scalings(X, a, L, U):
    """b_d = ln sum_v x_dv - ln sum_k exp(a_k + l_k . u_d)   (reference :141-179)."""
    return np.log(X.sum(axis=1)) - np.log(np.exp(a[:, None] + L @ U.T).sum(axis=0))


# ---- one embedding problem (reference :182-351): minimise
#      f(e) = -[ sum_i (o_i . e) aux_i - sum_i exp(s + s_i + o_i . e) - |e|^2 / (2 var) ]
def embedding_f(e, others, s, s_others, variance, aux_vec):
    sp = others @ e
    return -(float(sp @ aux_vec) - float(np.exp(s + s_others + sp).sum()) - float(e @ e) / (2 * variance))


def embedding_grad(e, others, s, s_others, variance, aux_vec):
    sp = others @ e
    return (np.exp(s + s_others + sp)[:, None] * others).sum(axis=0) - (aux_vec[:, None] * others).sum(axis=0) + e / variance


def embedding_hess(e, others, s, s_others, variance):
    sp = others @ e
    w = np.exp(s + s_others + sp)
    return (w[:, None, None] * others[:, :, None] * others[:, None, :]).sum(axis=0) + np.eye(len(e)) / variance


def snap(e):
    """Components with 0 < |e_j| < EPSILON are moved to +-EPSILON (reference :408-409)."""
    e = np.array(e, dtype=np.float64)
    e[(0 < e) & (e < EPSILON)] = EPSILON
    e[(-EPSILON < e) & (e < 0)] = -EPSILON
    return e


def update_embedding(e0, others, s, s_others, variance, aux_vec, maxiter=None, solver="scipy"):
    """Reference :354-410.  ``solver='own'`` uses the restated Newton-CG below."""
    f = lambda e: embedding_f(e, others, s, s_others, variance, aux_vec)  # noqa: E731
    g = lambda e: embedding_grad(e, others, s, s_others, variance, aux_vec)  # noqa: E731
    h = lambda e: embedding_hess(e, others, s, s_others, variance)  # noqa: E731
    if solver == "scipy":
        opts = {} if maxiter is None else {"maxiter": maxiter}
        x = optimize.minimize(fun=f, x0=np.array(e0, dtype=np.float64), method="Newton-CG", jac=g, hess=h, options=opts).x
    else:
        x = newton_cg(f, g, h, np.array(e0, dtype=np.float64), maxiter=maxiter)
    return snap(x)


def update_signature_embeddings(aux, a, b, L, U, variance, solver="scipy"):
    """k independent problems over the samples (reference corrnmf_det.py:88-113)."""
    return np.stack([update_embedding(L[j], U, a[j], b, variance, aux[j], None, solver) for j in range(L.shape[0])])


def update_sample_embeddings(aux, a, b, L, U, variance, solver="scipy"):
    """D independent problems over the signatures, 3 Newton iterations each (reference corrnmf_det.py:115-141)."""
    return np.stack([update_embedding(U[d], L, b[d], a, variance, aux[:, d], 3, solver) for d in range(U.shape[0])])


def update_variance(L, U):
    """Reference corrnmf_det.py:60-69."""
    return float(np.clip(np.mean(np.concatenate([L, U]) ** 2), EPSILON, None))


def update_parameters(X, W, a, b, L, U, variance, n_given_signatures=0, given=(), solver="scipy"):
    """One iteration in the reference's order (corrnmf_det.py:157-169); the signatures (and the ELBO) use the
    exposures computed BEFORE the scaling / embedding updates.  Returns the new state and those
    exposures."""
    if "sample_scalings" not in given:
        b = update_sample_scalings(X, a, L, U)
    H = compute_exposures(a, b, L, U)
    aux = compute_aux(X, W, H)
    if "signature_scalings" not in given:
        a = update_signature_scalings(aux, b, L, U)
    if "signature_embeddings" not in given:
        L = update_signature_embeddings(aux, a, b, L, U, variance, solver)
    if "sample_embeddings" not in given:
        U = update_sample_embeddings(aux, a, b, L, U, variance, solver)
    if "variance" not in given:
        variance = update_variance(L, U)
    W = klnmf.update_W(X.T, W.T, H.T, None, n_given_signatures).T
    return W, a, b, L, U, variance, H


# ---- restatement of SciPy's Newton-CG (scipy/optimize/_optimize.py::_minimize_newtoncg, defaults xtol=1e-5,
#      c1=1e-4, c2=0.9) with the DCSRCH line search (scipy/optimize/_dcsrch.py, MINPACK-2; amax=50, amin=1e-8,
#      xtol=1e-14, at most 100 trial steps).  Failure of DCSRCH ends the minimisation at the current point (SciPy
#      would first try its second line search).
def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = np.sign(dp) * np.sign(dx)
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) <= abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0 and gamma != 0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            stpf = min(stp + 0.66 * (sty - stp), stpf) if stp > stx else max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = min(max(stpf, stpmin), stpmax)
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * math.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


def dcsrch(phi, derphi, alpha1, phi0, derphi0, ftol=1e-4, gtol=0.9, xtol=1e-14, stpmin=1e-8, stpmax=50.0, maxiter=100):
    """Returns (stp or None, phi(stp), derphi-evaluated gradient is the caller's business)."""
    if alpha1 < stpmin or alpha1 > stpmax or derphi0 >= 0:
        return None, phi0
    brackt, stage = False, 1
    finit, ginit = phi0, derphi0
    gtest = ftol * ginit
    width = stpmax - stpmin
    width1 = width / 0.5
    stx = sty = 0.0
    fx = fy = finit
    gx = gy = ginit
    stmin, stmax = 0.0, alpha1 + 4.0 * alpha1
    stp = alpha1
    for _ in range(maxiter - 1):  # the first of SciPy's 100 _iterate calls only initialises
        f, g = phi(stp), derphi(stp)
        ftest = finit + stp * gtest
        if stage == 1 and f <= ftest and g >= 0:
            stage = 2
        warn = False
        if brackt and (stp <= stmin or stp >= stmax):
            warn = True
        if brackt and stmax - stmin <= xtol * stmax:
            warn = True
        if stp == stpmax and f <= ftest and g <= gtest:
            warn = True
        if stp == stpmin and (f > ftest or g >= gtest):
            warn = True
        if f <= ftest and abs(g) <= gtol * -ginit:
            return stp, f  # convergence (takes precedence over the warnings, as in the reference code)
        if warn:
            return None, f
        if stage == 1 and f <= fx and f > ftest:
            fm, fxm, fym = f - stp * gtest, fx - stx * gtest, fy - sty * gtest
            gm, gxm, gym = g - gtest, gx - gtest, gy - gtest
            stx, fxm, gxm, sty, fym, gym, stp, brackt = _dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax)
            fx, fy = fxm + stx * gtest, fym + sty * gtest
            gx, gy = gxm + gtest, gym + gtest
        else:
            stx, fx, gx, sty, fy, gy, stp, brackt = _dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax)
        if brackt:
            if abs(sty - stx) >= 0.66 * width1:
                stp = stx + 0.5 * (sty - stx)
            width1 = width
            width = abs(sty - stx)
        if brackt:
            stmin, stmax = min(stx, sty), max(stx, sty)
        else:
            stmin, stmax = stp + 1.1 * (stp - stx), stp + 4.0 * (stp - stx)
        stp = min(max(stp, stpmin), stpmax)
        if (brackt and (stp <= stmin or stp >= stmax)) or (brackt and stmax - stmin <= xtol * stmax):
            stp = stx
        if not math.isfinite(stp):
            return None, f
    return None, phi0


def newton_cg(f, grad, hess, x0, maxiter=None, avextol=1e-5):
    x = np.array(x0, dtype=np.float64)
    n = len(x)
    if maxiter is None:
        maxiter = 200 * n
    cg_maxiter = 20 * n
    xtol = n * avextol
    eps64 = np.finfo(np.float64).eps
    old_fval, old_old_fval = f(x), None
    update_l1 = np.finfo(float).max
    k = 0
    while update_l1 > xtol:
        if k >= maxiter:
            break
        b = -grad(x)
        maggrad = np.abs(b).sum()
        termcond = min(0.5, math.sqrt(maggrad)) * maggrad
        xsupi = np.zeros(n)
        ri = -b
        psupi = -ri
        i = 0
        dri0 = float(ri @ ri)
        A = hess(x)
        failed = True
        for _ in range(cg_maxiter):
            if np.abs(ri).sum() <= termcond:
                failed = False
                break
            Ap = A @ psupi
            curv = float(psupi @ Ap)
            if 0 <= curv <= 3 * eps64:
                failed = False
                break
            elif curv < 0:
                if i == 0:
                    xsupi = dri0 / (-curv) * b
                failed = False
                break
            alphai = dri0 / curv
            xsupi = xsupi + alphai * psupi
            ri = ri + alphai * Ap
            dri1 = float(ri @ ri)
            psupi = -ri + (dri1 / dri0) * psupi
            i += 1
            dri0 = dri1
        if failed:
            break  # "CG iterations didn't converge"
        pk, gfk = xsupi, -b
        derphi0 = float(gfk @ pk)
        if old_old_fval is not None and derphi0 != 0:
            alpha1 = min(1.0, 1.01 * 2 * (old_fval - old_old_fval) / derphi0)
            if alpha1 < 0:
                alpha1 = 1.0
        else:
            alpha1 = 1.0
        stp, fval = dcsrch(lambda s: f(x + s * pk), lambda s: float(grad(x + s * pk) @ pk), alpha1, old_fval, derphi0)
        if stp is None:
            break  # line search failed
        old_old_fval, old_fval = old_fval, fval
        update = stp * pk
        x = x + update
        k += 1
        update_l1 = np.abs(update).sum()
    return x
