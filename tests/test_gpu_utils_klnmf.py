"""
GPU parity of the fused KL-NMF pass + W epilogue, through the C ABI.

Part 1 mirrors reference tests/test_utils_klnmf.py:48-196 on the reference's golden fixtures
(96 x 10, k in {1, 2}, zeros in the counts).  Part 2 compares with the oracle on seeded inputs
covering ragged shapes (D not a multiple of the tile, V in {96, 83, 32}, every k padding class),
per-sample weights, the l-half branch and given signatures.
Tolerances: float64 rtol 1e-11 (different summation order only); float32 rtol 3e-5.
"""

import numpy as np
import pytest
import torch
from conftest import golden_path, load_counts
from gpu_api import GpuKL

from oracle import EPSILON, klnmf

pytestmark = pytest.mark.gpu

U = "models/utils_klnmf"
TOL = {torch.float64: dict(rtol=1e-11, atol=1e-13), torch.float32: dict(rtol=3e-5, atol=1e-9)}


@pytest.fixture(params=[torch.float64, torch.float32], ids=["f64", "f32"])
def gpu(request):
    return GpuKL(request.param)


@pytest.fixture(params=[1, 2])
def k(request):
    return request.param


@pytest.fixture
def xwh(k):
    X = load_counts(U, "counts.csv").astype(float)
    return X, np.load(golden_path(U, f"W_nsigs{k}.npy")), np.load(golden_path(U, f"H_nsigs{k}.npy"))


def g(name):
    return np.load(golden_path(U, name))


# ---- part 1: the reference's golden vectors ----------------------------------------------


def test_kl_divergence(gpu, xwh, k):
    assert np.allclose(gpu.kl_divergence(*xwh), g(f"kl_divergence_nsigs{k}.npy"))
    w = 2 * np.ones(xwh[0].shape[1])
    assert np.allclose(gpu.kl_divergence(*xwh, w), 2 * g(f"kl_divergence_nsigs{k}.npy"))


def test_samplewise_kl_divergence(gpu, xwh, k):
    assert np.allclose(gpu.samplewise_kl_divergence(*xwh), g(f"samplewise_kl_divergence_nsigs{k}.npy"))


def test_poisson_llh(gpu, xwh, k):
    from scipy.special import gammaln

    got = gpu.poisson_llh_wo_factorial(*xwh) - gammaln(1 + xwh[0]).sum()
    assert np.allclose(got, g(f"poisson_llh_nsigs{k}.npy"))


def test_update_W(gpu, xwh, k):
    ref = g(f"W_updated_standard_nsigs{k}.npy")
    assert np.allclose(gpu.update_W(*xwh), ref)
    assert np.allclose(gpu.update_W(*xwh, 2 * np.ones(xwh[0].shape[1])), ref)


def test_given_signatures_update_W(gpu, xwh, k):
    X, W, H = xwh
    for n_given in range(1, k + 1):
        out = gpu.update_W(X, W.copy(), H, n_given_signatures=n_given)
        expect = W[:, :n_given] if gpu.dtype == torch.float64 else W[:, :n_given].astype(np.float32)
        assert np.array_equal(out[:, :n_given], expect)


def test_update_H(gpu, xwh, k):
    ref = g(f"H_updated_standard_nsigs{k}.npy")
    D = xwh[0].shape[1]
    assert np.allclose(gpu.update_H(*xwh), ref)
    assert np.allclose(gpu.update_H(*xwh, 2 * np.ones(D), np.zeros(D)), ref)


def test_update_WH(gpu, xwh, k):
    Wr, Hr = g(f"W_updated_joint_nsigs{k}.npy"), g(f"H_updated_joint_nsigs{k}.npy")
    D = xwh[0].shape[1]
    for args in [(), (2 * np.ones(D),), (2 * np.ones(D), np.zeros(D))]:
        Wn, Hn = gpu.update_WH(*xwh, *args)
        assert np.allclose(Wn, Wr) and np.allclose(Hn, Hr)


def test_given_signatures_update_WH(gpu, xwh, k):
    X, W, H = xwh
    for n_given in range(1, k + 1):
        Wn, _ = gpu.update_WH(X, W.copy(), H, n_given_signatures=n_given)
        expect = W[:, :n_given] if gpu.dtype == torch.float64 else W[:, :n_given].astype(np.float32)
        assert np.array_equal(Wn[:, :n_given], expect)


# ---- part 2: oracle on seeded, ragged inputs ----------------------------------------------


def _problem(V, D, k, seed, zeros=True):
    rng = np.random.default_rng(seed)
    W = rng.dirichlet(0.5 * np.ones(V), size=k).T.clip(EPSILON)
    H = rng.gamma(1.0, 200.0, size=(k, D)).clip(EPSILON)
    X = rng.poisson(W @ H).astype(float)
    if zeros:
        X[rng.random(X.shape) < 0.05] = 0.0
    return X, W, H, rng


SHAPES = [(96, 1, 1), (96, 191, 5), (96, 193, 8), (96, 1000, 9), (83, 777, 16), (32, 400, 17), (96, 2049, 20), (96, 500, 25), (96, 300, 32), (7, 50, 3)]


@pytest.mark.parametrize("V,D,k", SHAPES)
def test_pass_matches_oracle(gpu, V, D, k):
    X, W, H, rng = _problem(V, D, k, seed=V * 1000 + D + k)
    tol = TOL[gpu.dtype]
    wk = rng.uniform(0.5, 2.0, D)
    wl = rng.uniform(0.0, 5.0, D)
    assert np.isclose(gpu.kl_divergence(X, W, H), klnmf.kl_divergence(X, W, H), rtol=tol["rtol"])
    assert np.isclose(gpu.kl_divergence(X, W, H, wk, wl), klnmf.klnmf_objective(X, W, H, wk, wl), rtol=tol["rtol"])
    assert np.allclose(gpu.samplewise_kl_divergence(X, W, H), klnmf.samplewise_kl_divergence(X, W, H), rtol=tol["rtol"] * 10, atol=1e-4 if gpu.dtype == torch.float32 else 1e-9)
    Xc = X.clip(EPSILON)
    for n_given in sorted({0, 1, k}):
        Wn, Hn = gpu.update_WH(Xc, W, H, None, None, n_given)
        Wr, Hr = klnmf.update_WH(Xc, W, H, None, None, n_given)
        assert np.allclose(Wn, Wr, **tol) and np.allclose(Hn, Hr, **tol)
        Wn, Hn = gpu.update_WH(Xc, W, H, wk, wl, n_given)
        Wr, Hr = klnmf.update_WH(Xc, W, H, wk, wl, n_given)
        assert np.allclose(Wn, Wr, **tol)
        assert np.allclose(Hn, Hr, rtol=tol["rtol"] * 30, atol=tol["atol"] + (1e-3 if gpu.dtype == torch.float32 else 0))
        assert np.allclose(gpu.update_W(Xc, W, H, wk, n_given), klnmf.update_W(Xc, W, H, wk, n_given), **tol)
    assert np.allclose(gpu.update_H(Xc, W, H), klnmf.update_H(Xc, W, H), **tol)


def test_empty_shard_and_bad_arguments():
    from salamander_b200._device import PASS_OBJECTIVE, PASS_WNUM, Workspace

    dev = torch.device("cuda:0")
    ws = Workspace(96, 0, 3, torch.float64, dev)
    X = torch.zeros((0, 96), dtype=torch.float64, device=dev)
    H = torch.zeros((0, 3), dtype=torch.float64, device=dev)
    W = torch.full((3, 96), 1 / 96, dtype=torch.float64, device=dev)
    Wnum = torch.ones_like(W)
    obj = torch.ones(1, dtype=torch.float64, device=dev)
    ws.klnmf_pass(X, W, H, PASS_WNUM | PASS_OBJECTIVE, Wnum=Wnum, objective=obj)
    assert float(obj.item()) == 0.0 and float(Wnum.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        ws.klnmf_pass(X, W, H, PASS_WNUM)  # Wnum missing
    with pytest.raises(ValueError):
        ws.klnmf_pass(X, W.float(), H, PASS_OBJECTIVE, objective=obj)  # wrong dtype
    ws.close()
    with pytest.raises(NotImplementedError):
        Workspace(97, 10, 3, torch.float64, dev)
    with pytest.raises(NotImplementedError):
        Workspace(96, 10, 33, torch.float64, dev)


def test_shard_linearity_full_size():
    """Size-independent property at a large D: numerators / objectives of two row shards add up to
    the unsharded result (this is exactly what the multi-GPU all-reduce relies on)."""
    from salamander_b200._device import PASS_OBJECTIVE, PASS_WNUM, Workspace

    dev = torch.device("cuda:0")
    V, D, k = 96, 200_003, 20
    gen = torch.Generator(device=dev).manual_seed(5)
    W = torch.rand((k, V), generator=gen, device=dev, dtype=torch.float64)
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=gen, device=dev, dtype=torch.float64) * 500 + 1e-3
    X = torch.poisson(H @ W, generator=gen).clamp_min(EPSILON)
    out = []
    for lo, hi in [(0, D), (0, 77_777), (77_777, D)]:
        ws = Workspace(V, hi - lo, k, torch.float64, dev)
        Wnum = torch.empty_like(W)
        obj = torch.zeros(1, dtype=torch.float64, device=dev)
        ws.klnmf_pass(X[lo:hi].contiguous(), W, H[lo:hi].contiguous(), PASS_WNUM | PASS_OBJECTIVE, Wnum=Wnum, objective=obj)
        out.append((Wnum.clone(), obj.clone()))
        ws.close()
    assert torch.allclose(out[0][0], out[1][0] + out[2][0], rtol=1e-11)
    assert torch.allclose(out[0][1], out[1][1] + out[2][1], rtol=1e-12)
    # and against plain torch linear algebra (fp64) on the device
    A = X / (H @ W)
    assert torch.allclose(out[0][0], H.T @ A, rtol=1e-10)
