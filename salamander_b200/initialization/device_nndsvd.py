"""
NNDSVD initialisation with the heavy linear algebra on the device.

The reference calls scikit-learn's ``_initialize_nmf`` (initialization/methods.py:69-86), i.e. a *randomized* truncated SVD
of the (n_samples x 96) count matrix followed by the NNDSVD construction of Boutsidis & Gallopoulos (2008).  On a million
samples that host SVD takes seconds -- longer than the whole fit on a B200.  Here the truncated SVD is exact and almost
free: the 96 x 96 Gram matrix X^T X is accumulated on the device in float64 (one pass over X), its eigen-decomposition
(host, 96 x 96) gives the right singular vectors and singular values, and the left singular vectors U = X V / s and every
per-sample quantity of NNDSVD are formed on the device.  NNDSVD is invariant to the sign of a singular pair, so no sign
convention is needed.

Agreement with scikit-learn is limited by the *randomized* SVD's own approximation error, not by this code
(tests/test_gpu_init.py: entries agree to ~1e-6 relative on the PCAWG counts, where the spectrum decays fast).  It is
therefore opt-in for parity work: ``KLNMF(..., init_device=True)`` or ``"auto"`` (matrices of at least 2^22 entries).

PyTorch is used for the device linear algebra (cuBLAS GEMMs); this is initialisation, not the fitting path.
"""

from __future__ import annotations

import numpy as np
import torch

SK_EPS = 1e-6  # scikit-learn's threshold below which NNDSVD entries are set to zero
N_OVERSAMPLES = 10  # randomized_svd's default: it draws a (n_features, k + 10) normal matrix from the global RNG


def _upload_chunks(data_mat: np.ndarray, device: torch.device, rows: int = 1 << 18):
    """The matrix as float64 device tensors, uploaded once in row blocks (D x 96 doubles: 768 MB per million samples)."""
    out = []
    for lo in range(0, data_mat.shape[0], rows):
        out.append((lo, torch.from_numpy(np.ascontiguousarray(data_mat[lo : lo + rows])).to(device, non_blocking=True).to(torch.float64)))
    return out


def truncated_svd_gram(data_mat: np.ndarray, k: int, device: torch.device):
    """Top-k singular triplets of ``data_mat`` (D x V, V small): (U on the device [D][k] float64, s [k], Vt [k][V])."""
    D, V = data_mat.shape
    chunks = _upload_chunks(data_mat, device)
    gram = torch.zeros((V, V), dtype=torch.float64, device=device)
    for _, chunk in chunks:
        gram.addmm_(chunk.T, chunk)
    evals, evecs = np.linalg.eigh(gram.cpu().numpy())
    top = np.argsort(evals)[::-1][:k]
    s = np.sqrt(np.maximum(evals[top], 0.0))
    Vk = np.ascontiguousarray(evecs[:, top])  # [V][k]
    proj = torch.from_numpy(Vk / np.where(s > 0, s, 1.0)).to(device)
    U = torch.empty((D, k), dtype=torch.float64, device=device)
    for lo, chunk in chunks:
        torch.mm(chunk, proj, out=U[lo : lo + chunk.shape[0]])
    return U, s, Vk.T


def init_nndsvd_device(data_mat: np.ndarray, n_signatures: int, method: str = "nndsvd", seed: int | None = None, device=None):
    """Same contract as ``methods.init_nndsvd``: returns (signatures_mat (k, V), exposures_mat (D, k)) as host arrays."""
    if seed is not None:
        np.random.seed(seed)
    D, V = data_mat.shape
    k = int(n_signatures)
    if k > min(D, V):
        raise ValueError(f"init = '{method}' can only be used when n_components <= min(n_samples, n_features)")
    device = torch.device("cuda") if device is None else torch.device(device)
    # keep the global RNG stream where scikit-learn would leave it (the range finder's Gaussian test matrix)
    np.random.normal(size=(V, k + N_OVERSAMPLES))

    U, s, Vt = truncated_svd_gram(np.asarray(data_mat), k, device)

    # positive / negative parts of every singular pair, all components at once
    y_pos, y_neg = np.maximum(Vt, 0.0), np.maximum(-Vt, 0.0)
    yp, yn = np.linalg.norm(y_pos, axis=1), np.linalg.norm(y_neg, axis=1)
    x_pos, x_neg = U.clamp_min(0.0), (-U).clamp_min(0.0)
    xp, xn = torch.linalg.vector_norm(x_pos, dim=0).cpu().numpy(), torch.linalg.vector_norm(x_neg, dim=0).cpu().numpy()
    use_pos = xp * yp > xn * yn
    sigma = np.where(use_pos, xp * yp, xn * yn)
    lam = np.sqrt(s * sigma)
    x_nrm, y_nrm = np.where(use_pos, xp, xn), np.where(use_pos, yp, yn)
    with np.errstate(divide="ignore", invalid="ignore"):
        col_scale = lam / x_nrm
        row_scale = lam / y_nrm
    pick = torch.from_numpy(use_pos).to(device)
    expo = torch.where(pick[None, :], x_pos, x_neg) * torch.from_numpy(col_scale).to(device)[None, :]
    sigs = np.where(use_pos[:, None], y_pos, y_neg) * row_scale[:, None]
    # the leading pair is used as it is (non-negative up to sign by Perron-Frobenius)
    expo[:, 0] = np.sqrt(s[0]) * U[:, 0].abs()
    sigs[0] = np.sqrt(s[0]) * np.abs(Vt[0])
    expo = torch.where(expo < SK_EPS, torch.zeros_like(expo), expo)
    sigs[sigs < SK_EPS] = 0.0

    if method != "nndsvd":
        avg = float(np.mean(data_mat))
        if method == "nndsvda":
            expo = torch.where(expo == 0, torch.full_like(expo, avg), expo)
            sigs[sigs == 0] = avg
        elif method == "nndsvdar":
            # same draws, in the same order, as scikit-learn: first the zeros of the exposure factor, then the signatures'
            zeros_e = (expo == 0).cpu().numpy()
            fill = np.abs(avg * np.random.standard_normal(size=int(zeros_e.sum())) / 100)
            expo_host = expo.cpu().numpy()
            expo_host[zeros_e] = fill
            zeros_s = sigs == 0
            sigs[zeros_s] = np.abs(avg * np.random.standard_normal(size=int(zeros_s.sum())) / 100)
            return sigs, expo_host
        else:
            raise ValueError(f"unknown NNDSVD variant {method!r}")
    return sigs, expo.cpu().numpy()


def init_random_device(data_mat: np.ndarray, n_signatures: int, seed: int | None = None, device=None, _row_totals=None):
    """``init_random`` (reference initialization/methods.py:89-109) with the D x k exposure draws on the device.

    Signatures: the reference's own draw, ``np.random.dirichlet(1_V, size=k)`` from the global numpy RNG (so W0 is the
    reference's for a given seed).  Exposures: sample total x Dirichlet(1_k), drawn as normalised Exp(1) variates by a
    torch generator on the device that is seeded from the same numpy stream -- the same distribution and reproducible
    for a seed, but NOT numpy's draws (its legacy gamma sampler cannot be reproduced cheaply on a GPU).  On 100k samples
    and k = 30 the host draws take ~80 ms, longer than 500 updates of the fit; this takes ~1 ms.  Opt-in (``init_device``).
    """
    if seed is not None:
        np.random.seed(seed)
    D, V = data_mat.shape
    k = int(n_signatures)
    device = torch.device("cuda") if device is None else torch.device(device)
    sigs = np.random.dirichlet(np.ones(V), size=k)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(np.random.randint(0, 2**31 - 1)))
    resident = _row_totals is not None  # (a sweep's resident counts: totals already on the device, exposures stay there)
    totals = _row_totals if resident else torch.from_numpy(np.asarray(data_mat).sum(axis=1, dtype=np.float64)).to(device)
    e = torch.empty((D, k), dtype=torch.float64, device=device)
    e.exponential_(1.0, generator=gen)
    e.mul_((totals / e.sum(dim=1))[:, None])
    return sigs, (e if resident else e.cpu().numpy())
