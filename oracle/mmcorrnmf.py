"""
Oracle (test infrastructure): multimodal correlated NMF in numpy float64 -- several CorrNMF models (one per
modality) sharing the sample embeddings.  Restates the numerics of reference models/mmcorrnmf.py: ELBO :168-194,
per-modality updates :247-396, joint sample-embedding update with a VECTOR scaling :398-428, pooled variance
:305-317, update order :443-453.  The per-modality pieces are oracle/corrnmf.py.

A modality is a dict with keys X (D, V), W (k, V), a (k,), b (D,), L (k, m) and, after ``update_parameters``,
H (D, k): the exposures computed before the scaling / embedding updates, which the signature update and the ELBO use.
"""

from __future__ import annotations

import numpy as np

from . import EPSILON, corrnmf, klnmf


def compute_exposures(mods, U):
    for md in mods:
        md["H"] = corrnmf.compute_exposures(md["a"], md["b"], md["L"], U)


def compute_auxs(mods):
    return [corrnmf.compute_aux(md["X"], md["W"], md["H"]) for md in mods]


def elbo(mods, U, variance):
    """Reference mmcorrnmf.py:168-194: modality ELBOs without the sample-embedding prior, which is added once."""
    val = sum(corrnmf.elbo(md["X"], md["W"], md["H"], md["L"], U, variance, penalize_sample_embeddings=False) for md in mods)
    D, m = U.shape
    val -= 0.5 * m * D * np.log(2 * np.pi * variance)
    val -= np.sum(U**2) / (2 * variance)
    return float(val)


def update_sample_embeddings(mods, auxs, U, variance, solver="scipy"):
    """Reference :398-428: others = all signature embeddings, scalings_other = all signature scalings, and the
    sample's scaling is a vector: its scaling in modality j repeated k_j times."""
    L_all = np.concatenate([md["L"] for md in mods])
    a_all = np.concatenate([md["a"] for md in mods])
    aux_all = np.concatenate(auxs)
    out = np.empty_like(U)
    for d in range(U.shape[0]):
        scal = np.concatenate([np.repeat(md["b"][d], md["L"].shape[0]) for md in mods])
        out[d] = corrnmf.update_embedding(U[d], L_all, scal, a_all, variance, aux_all[:, d], 3, solver)
    return out


def update_variance(mods, U):
    emb = np.concatenate([md["L"] for md in mods] + [U])
    return float(np.clip(np.mean(emb**2), EPSILON, None))


def update_parameters(mods, U, variance, solver="scipy"):
    """One iteration in the reference's order (:443-453).  Mutates ``mods``; returns (U, variance)."""
    for md in mods:
        md["b"] = corrnmf.update_sample_scalings(md["X"], md["a"], md["L"], U)
    compute_exposures(mods, U)
    auxs = compute_auxs(mods)
    for md, aux in zip(mods, auxs):
        md["a"] = corrnmf.update_signature_scalings(aux, md["b"], md["L"], U)
    for md, aux in zip(mods, auxs):
        md["L"] = corrnmf.update_signature_embeddings(aux, md["a"], md["b"], md["L"], U, variance, solver)
    U = update_sample_embeddings(mods, auxs, U, variance, solver)
    variance = update_variance(mods, U)
    for md in mods:
        md["W"] = klnmf.update_W(md["X"].T, md["W"].T, md["H"].T).T
    return U, variance
