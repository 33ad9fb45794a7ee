"""
GPU parity of CorrNMFDet through the public API, mirroring reference tests/test_corrnmf.py:108-245 on the
reference's golden fixtures (tests/golden/models/corrnmf), plus several whole iterations against the oracle
(whose embedding solver is scipy's Newton-CG, as in the reference) on the PCAWG counts.
"""

import os

import ctypes as C

import numpy as np
import pandas as pd
import pytest
import torch
from conftest import ROOT, golden_path

import salamander_b200 as sal
from salamander_b200 import AnnData

from oracle import corrnmf as oracle

pytestmark = pytest.mark.gpu
P = "models/corrnmf"


def counts():
    return pd.read_csv(golden_path(P, "counts.csv"), index_col=0).T


@pytest.fixture(params=[1, 2])
def n(request):
    return request.param


@pytest.fixture(params=["float64"])
def dtype(request):
    return request.param


def ld(name, n):
    return np.load(golden_path(P, f"{name}_nsigs{n}_dim{n}.npy"))


@pytest.fixture
def model_init(n, dtype):
    adata = AnnData(counts())
    adata.obs["scalings"] = ld("sample_scalings_init", n)
    adata.obsm["embeddings"] = ld("sample_embeddings_init", n).T
    asig = AnnData(ld("signatures_mat_init", n).T)
    asig.var_names = adata.var_names
    asig.obs["scalings"] = ld("signature_scalings_init", n)
    asig.obsm["embeddings"] = ld("signature_embeddings_init", n).T
    model = sal.models.CorrNMFDet(n_signatures=n, dim_embeddings=n, dtype=dtype)
    model.adata = adata
    model.asignatures = asig
    model.compute_exposures()
    model.variance = float(ld("variance_init", n))
    return model


def test_objective_function(model_init, n):
    assert np.allclose(model_init.objective_function(), ld("objective_init", n))


def test_compute_aux(model_init, n):
    assert np.allclose(model_init._compute_aux(), ld("aux", n))


def test_update_signatures(model_init, n):
    model_init.update_signatures()
    assert np.allclose(model_init.asignatures.X, ld("signatures_mat_updated", n).T)


def test_update_signature_scalings(model_init, n):
    model_init.update_signature_scalings(ld("aux", n))
    assert np.allclose(model_init.asignatures.obs["scalings"].values, ld("signature_scalings_updated", n))


def test_update_sample_scalings(model_init, n):
    model_init.update_sample_scalings()
    assert np.allclose(model_init.adata.obs["scalings"].values, ld("sample_scalings_updated", n))


def test_update_signature_embeddings(model_init, n):
    model_init.update_signature_embeddings(ld("aux", n))
    assert np.allclose(model_init.asignatures.obsm["embeddings"], ld("signature_embeddings_updated", n).T)


def test_update_sample_embeddings(model_init, n):
    model_init.update_sample_embeddings(ld("aux", n))
    assert np.allclose(model_init.adata.obsm["embeddings"], ld("sample_embeddings_updated", n).T)


def test_update_variance(model_init, n):
    model_init.update_variance()
    assert np.allclose(model_init.variance, ld("variance_updated", n))


@pytest.mark.parametrize("k,m", [(1, 1), (2, 1), (2, 2)])
def test_given_parameters_stay_fixed(k, m):
    adata = AnnData(counts())
    rng = np.random.default_rng(0)

    def model():
        return sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, min_iterations=3, max_iterations=3, init_method="random")

    for n_given in range(1, k + 1):
        given = adata[:n_given, :].copy()
        given.X = given.X / np.sum(given.X, axis=1, keepdims=True)
        mod = model()
        mod.fit(adata.copy(), given_parameters={"asignatures": given}, init_kwargs={"seed": 1})
        assert np.allclose(given.X, mod.asignatures.X[:n_given, :])
    cases = {
        "signature_scalings": rng.uniform(size=k),
        "sample_scalings": rng.uniform(size=adata.n_obs),
        "signature_embeddings": rng.uniform(size=(k, m)),
        "sample_embeddings": rng.uniform(size=(adata.n_obs, m)),
        "variance": 3,
    }
    for key, val in cases.items():
        mod = model()
        mod.fit(adata.copy(), given_parameters={key: val}, init_kwargs={"seed": 1})
        got = {
            "signature_scalings": lambda: mod.asignatures.obs["scalings"].values,
            "sample_scalings": lambda: mod.adata.obs["scalings"].values,
            "signature_embeddings": lambda: mod.asignatures.obsm["embeddings"],
            "sample_embeddings": lambda: mod.adata.obsm["embeddings"],
            "variance": lambda: mod.variance,
        }[key]()
        assert np.allclose(val, got), key
    with pytest.raises(ValueError):
        model().fit(adata.copy(), given_parameters={"variance": -1.0})
    with pytest.raises(ValueError):
        sal.models.CorrNMFDet(n_signatures=2, init_method="custom").fit(adata.copy())


def test_iterations_match_the_oracle_on_pcawg():
    """5 whole iterations (k = 4, dim 3) from the same start: every parameter and the ELBO history."""
    cnt = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T
    k, m, n_iter = 4, 3, 5
    model = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, init_method="random", min_iterations=n_iter, max_iterations=n_iter,
                                  conv_test_freq=1)
    adata = AnnData(cnt)
    model._setup_adata(adata)
    np.random.seed(3)
    model._initialize(None, {"seed": 3})
    X = np.asarray(adata.X, dtype=float)
    W = np.array(model.asignatures.X)
    a, b = np.array(model.asignatures.obs["scalings"].values, dtype=float), np.array(adata.obs["scalings"].values, dtype=float)
    L, U = np.array(model.asignatures.obsm["embeddings"]), np.array(adata.obsm["embeddings"])
    var = float(model.variance)
    hist_ref = []
    for _ in range(n_iter):
        W, a, b, L, U, var, H = oracle.update_parameters(X, W, a, b, L, U, var)
        hist_ref.append(oracle.elbo(X, W, H, L, U, var))
    # run the device model from the very same start
    with model._resident():
        model._in_fit = True
        hist = []
        for _ in range(n_iter):
            model._update_parameters(None)
            hist.append(model.objective_function())
        model._in_fit = False
    assert np.allclose(hist, hist_ref, rtol=1e-8), (hist, hist_ref)
    assert np.allclose(model.asignatures.X, W, rtol=1e-6, atol=1e-12)
    assert np.allclose(model.asignatures.obs["scalings"].values, a, rtol=1e-6)
    assert np.allclose(model.adata.obs["scalings"].values, b, rtol=1e-6)
    assert np.allclose(model.asignatures.obsm["embeddings"], L, rtol=1e-5, atol=1e-8)
    assert np.allclose(model.adata.obsm["embeddings"], U, rtol=1e-5, atol=1e-7)
    assert np.isclose(model.variance, var, rtol=1e-7)


def test_iterations_match_the_oracle_on_synthetic_counts():
    """6 whole iterations (k = 5, dim 4) on 2,000 synthetic samples of the bench generator: larger counts and far more
    samples per signature problem than PCAWG, so the line searches and the CG inner loops take more varied paths."""
    import bench

    k, m, n_iter, D = 5, 4, 6, 2000
    X0 = bench.synth_rows(0, D, k).astype(np.float64)
    model = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, init_method="random", min_iterations=n_iter, max_iterations=n_iter,
                                  conv_test_freq=1)
    adata = AnnData(X0)
    model._setup_adata(adata)
    np.random.seed(0)
    model._initialize(None, {"seed": 0})
    X = np.asarray(adata.X, dtype=float)
    W = np.array(model.asignatures.X)
    a, b = np.array(model.asignatures.obs["scalings"].values, dtype=float), np.array(adata.obs["scalings"].values, dtype=float)
    L, U = np.array(model.asignatures.obsm["embeddings"]), np.array(adata.obsm["embeddings"])
    var = float(model.variance)
    hist_ref = []
    for _ in range(n_iter):
        W, a, b, L, U, var, H = oracle.update_parameters(X, W, a, b, L, U, var)
        hist_ref.append(oracle.elbo(X, W, H, L, U, var))
    with model._resident():
        model._in_fit = True
        hist = []
        for _ in range(n_iter):
            model._update_parameters(None)
            hist.append(model.objective_function())
        model._in_fit = False
    assert np.allclose(hist, hist_ref, rtol=1e-8), (hist, hist_ref)
    assert np.allclose(model.asignatures.obsm["embeddings"], L, rtol=1e-5, atol=1e-8)
    assert np.allclose(model.adata.obsm["embeddings"], U, rtol=1e-5, atol=1e-7)
    assert np.isclose(model.variance, var, rtol=1e-7)


@pytest.mark.parametrize("tag", ["corrnmf_pcawg_k4_dim3_seed3", "corrnmf_pcawg_k6_dim2_seed8"])
def test_fit_matches_the_live_reference_trajectory(tag):
    """``CorrNMFDet.fit`` against whole iterations of the LIVE reference (tests/golden/trajectories, written by
    oracle/make_golden.py::corrnmf_case in the build container): same seed, same start, ELBO after every iteration within
    1e-8, final parameters within the Newton-CG tolerances."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "trajectories", f"{tag}.npz"))
    k, dim, seed, n_iter = int(z["k"]), int(z["dim"]), int(z["seed"]), int(z["n_iter"])
    cnt = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T
    model = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=dim, init_method="random", min_iterations=n_iter, max_iterations=n_iter,
                                  conv_test_freq=1)
    np.random.seed(seed)
    model.fit(AnnData(cnt), init_kwargs={"seed": seed})
    hist = np.array(model.history["objective_function"])
    assert hist.shape == z["history"].shape
    assert np.allclose(hist, z["history"], rtol=1e-8, atol=0), (hist, z["history"])
    assert np.allclose(model.asignatures.X, z["W"], rtol=1e-6, atol=1e-12)
    assert np.allclose(model.asignatures.obs["scalings"].values, z["a"], rtol=1e-6)
    assert np.allclose(model.adata.obs["scalings"].values, z["b"], rtol=1e-6)
    assert np.allclose(model.asignatures.obsm["embeddings"], z["L"], rtol=1e-5, atol=1e-8)
    assert np.allclose(model.adata.obsm["embeddings"], z["U"], rtol=1e-5, atol=1e-7)
    assert np.isclose(model.variance, float(z["var"]), rtol=1e-7)


@pytest.mark.parametrize("D,k,m", [(192, 3, 2), (5000, 5, 4), (40_000, 4, 3), (70_001, 2, 6)])
def test_signature_embeddings_exchange_with_emulated_ranks(D, k, m):
    """The multi-GPU protocol of the signature-embedding solver (sal_corrnmf_signature_embeddings_p2p: every rank sweeps its own
    samples, the totals of each evaluation travel as tagged words and are summed in rank order) with two ranks emulated on ONE
    GPU: both ranks must arrive at bit-identical embeddings, equal to the single-rank solve to rounding, launch after launch
    on the same receive buffers."""
    from salamander_b200 import _lib
    from salamander_b200._device import Workspace, corrnmf_signature_embeddings_emulated

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(D + k)
    tens = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    lib = _lib.load()
    n_words = int(lib.sal_corrnmf_sig_exchange_bytes(k, 2)) // 16
    recv = [torch.zeros((n_words, 4), dtype=torch.int32, device=dev) for _ in range(2)]
    table = [torch.tensor([recv[0].data_ptr(), recv[1].data_ptr()], dtype=torch.int64, device=dev) for _ in range(2)]
    cut = D // 3  # unequal shards
    for launch in range(1, 4):
        U = rng.normal(size=(D, m)) * 0.5
        b = rng.normal(size=D) * 0.3 + 2.0
        a = rng.normal(size=k) * 0.2
        L0 = rng.normal(size=(k, m)) * 0.5
        H = np.exp(a[None, :] + b[:, None] + U @ L0.T)
        aux = H * rng.uniform(0.5, 1.5, size=(D, k))
        var = 0.7 + 0.1 * launch
        ws_all = Workspace(96, D, k, torch.float64, dev)
        L_one = tens(L0)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.sal_corrnmf_signature_embeddings(ws_all._h, *(C.c_void_p(t.data_ptr()) for t in (tens(aux), tens(a), tens(b), L_one, tens(U))),
                                                        m, var, stream), "sal_corrnmf_signature_embeddings")
        shards = [slice(0, cut), slice(cut, D)]
        wss = [Workspace(96, sl.stop - sl.start, k, torch.float64, dev) for sl in shards]
        Ls = [tens(L0), tens(L0)]
        keep = [[tens(aux[sl]) for sl in shards], [tens(a), tens(a)], [tens(b[sl]) for sl in shards], [tens(U[sl]) for sl in shards]]
        corrnmf_signature_embeddings_emulated(wss, keep[0], keep[1], keep[2], Ls, keep[3], m, var, table, launch)
        torch.cuda.synchronize()
        assert torch.equal(Ls[0], Ls[1]), "emulated ranks must hold bit-identical signature embeddings"
        assert not torch.equal(Ls[0], tens(L0))
        assert np.allclose(Ls[0].cpu().numpy(), L_one.cpu().numpy(), rtol=1e-7, atol=1e-10), (Ls[0], L_one)
        for w in wss + [ws_all]:
            w.close()


@pytest.mark.parametrize("n", [1, 10, 483, 4096])
def test_peer_memory_allreduce_with_emulated_ranks(n):
    """sal_p2p_allreduce_f64 (the few-KB sums of a sharded CorrNMF iteration) with two ranks emulated on one GPU: both end with
    the same bits, equal to the rank-ordered sum, call after call on the same receive buffers (tags only ever increase)."""
    from salamander_b200 import _lib
    from salamander_b200._device import Workspace

    dev = torch.device("cuda", 0)
    lib = _lib.load()
    assert n <= int(lib.sal_p2p_allreduce_max_values())
    ws = Workspace(96, 8, 2, torch.float64, dev)
    n_words = int(lib.sal_p2p_allreduce_bytes(2)) // 16
    recv = [torch.zeros((n_words, 4), dtype=torch.int32, device=dev) for _ in range(2)]
    tables = [torch.tensor([recv[0].data_ptr(), recv[1].data_ptr()], dtype=torch.int64, device=dev) for _ in range(2)]
    gen = torch.Generator(device=dev).manual_seed(n)
    for launch in range(1, 6):
        a = torch.randn(n, dtype=torch.float64, device=dev, generator=gen) * 10.0 ** launch
        b = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
        want = a + b  # rank 0's contribution first
        bufs = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())
        tabs = (C.c_void_p * 2)(tables[0].data_ptr(), tables[1].data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.sal_p2p_allreduce_f64_emulated(ws._h, bufs, n, tabs, 2, launch, stream), "sal_p2p_allreduce_f64_emulated")
        torch.cuda.synchronize()
        assert torch.equal(a, b) and torch.equal(a, want)
    ws.close()
