"""
Host-side initial guesses for the signature / exposure matrices.

Runs ONCE per fit on the host (outside the kernels' scope, needed so a
fit starts from the same W0, H0 as the reference).  Same method names, arguments and use
of the legacy global numpy RNG as reference initialization/methods.py:27-135 so that a
given ``seed`` reproduces the reference's draws exactly.

All functions take ``data_mat`` of shape (n_samples, n_features) and return
``(signatures_mat (k, V), exposures_mat (D, k))``.
"""

from __future__ import annotations

from typing import Literal, get_args

import numpy as np

from ..utils import shape_checker, type_checker

_Init_methods = Literal["custom", "flat", "nndsvd", "nndsvda", "nndsvdar", "random", "separableNMF"]
_INIT_METHODS = get_args(_Init_methods)


def init_custom(data_mat, n_signatures, signatures_mat, exposures_mat):
    """User-supplied matrices, only validated (reference methods.py:27-55)."""
    type_checker("signatures_mat", signatures_mat, np.ndarray)
    type_checker("exposures_mat", exposures_mat, np.ndarray)
    D, V = data_mat.shape
    shape_checker("signatures_mat", signatures_mat, (n_signatures, V))
    shape_checker("exposures_mat", exposures_mat, (D, n_signatures))
    return signatures_mat, exposures_mat


def init_flat(data_mat, n_signatures):
    """Uniform signatures; every exposure = sample total / k (reference methods.py:58-66)."""
    D, V = data_mat.shape
    sigs = np.full((n_signatures, V), 1.0 / V)
    per_sample = data_mat.sum(axis=1) / n_signatures
    expo = np.repeat(per_sample[:, None], n_signatures, axis=1)
    return sigs, expo


def init_nndsvd(data_mat, n_signatures, method="nndsvd", seed=None):
    """scikit-learn's NNDSVD family (reference methods.py:69-86); seeds the GLOBAL numpy RNG."""
    from sklearn.decomposition import _nmf as sknmf

    if seed is not None:
        np.random.seed(seed)
    expo, sigs = sknmf._initialize_nmf(data_mat, n_signatures, init=method)  # pylint: disable=protected-access
    return sigs, expo


def init_random(data_mat, n_signatures, seed=None):
    """Signatures ~ Dirichlet(1_V); exposures = sample total x Dirichlet(1_k) (reference methods.py:89-109)."""
    if seed is not None:
        np.random.seed(seed)
    D, V = data_mat.shape
    sigs = np.random.dirichlet(np.ones(V), size=n_signatures)
    totals = data_mat.sum(axis=1)
    expo = totals[:, None] * np.random.dirichlet(np.ones(n_signatures), size=D)
    return sigs, expo


def init_separableNMF(data_mat, n_signatures, seed=None):
    """Successive projection (Gillis & Vavasis 2013, Alg. 1, f = ||.||^2) picks k samples as signatures;
    exposures as in ``init_random`` (reference methods.py:112-135)."""
    picked = np.empty(n_signatures, dtype=int)
    R = data_mat.T / data_mat.T.sum(axis=0)
    n_feat = R.shape[0]
    for j in range(n_signatures):
        norms = (R**2).sum(axis=0)
        best = int(np.argmax(norms))
        u = R[:, best]
        R = (np.identity(n_feat) - np.outer(u, u) / norms[best]) @ R
        picked[j] = best
    sigs = data_mat[picked, :].astype(float)
    sigs /= sigs.sum(axis=1)[:, None]
    _, expo = init_random(data_mat, n_signatures, seed=seed)
    return sigs, expo
