"""
Pins the CorrNMF oracle (oracle/corrnmf.py) against the reference's own golden fixtures
(tests/golden/models/corrnmf = reference tests/test_data/models/corrnmf, used by reference
tests/test_corrnmf.py:108-175) and pins the restated Newton-CG against scipy's.  CPU only.
"""

import numpy as np
import pytest
from conftest import golden_path, load_counts

from oracle import corrnmf, klnmf

P = "models/corrnmf"


@pytest.fixture(params=[1, 2])
def fx(request):
    n = request.param
    suf = f"nsigs{n}_dim{n}.npy"
    ld = lambda name: np.load(golden_path(P, f"{name}_{suf}"))  # noqa: E731
    X = load_counts(P, "counts.csv").T.astype(float)  # (D, V)
    d = dict(
        X=X,
        W=ld("signatures_mat_init").T,
        a=ld("signature_scalings_init"),
        b=ld("sample_scalings_init"),
        L=ld("signature_embeddings_init").T,
        U=ld("sample_embeddings_init").T,
        var=float(ld("variance_init")),
        aux=ld("aux"),
        ld=ld,
    )
    d["H"] = corrnmf.compute_exposures(d["a"], d["b"], d["L"], d["U"])
    return d


def test_elbo(fx):
    assert np.allclose(corrnmf.elbo(fx["X"], fx["W"], fx["H"], fx["L"], fx["U"], fx["var"]), fx["ld"]("objective_init"))


def test_aux_and_signatures(fx):
    assert np.allclose(corrnmf.compute_aux(fx["X"], fx["W"], fx["H"]), fx["aux"])
    W = klnmf.update_W(fx["X"].T, fx["W"].T, fx["H"].T).T
    assert np.allclose(W, fx["ld"]("signatures_mat_updated").T)


def test_scalings_and_variance(fx):
    assert np.allclose(corrnmf.update_signature_scalings(fx["aux"], fx["b"], fx["L"], fx["U"]), fx["ld"]("signature_scalings_updated"))
    assert np.allclose(corrnmf.update_sample_scalings(fx["X"], fx["a"], fx["L"], fx["U"]), fx["ld"]("sample_scalings_updated"))
    assert np.allclose(corrnmf.update_variance(fx["L"], fx["U"]), fx["ld"]("variance_updated"))


@pytest.mark.parametrize("solver", ["scipy", "own"])
def test_embedding_updates(fx, solver):
    Ln = corrnmf.update_signature_embeddings(fx["aux"], fx["a"], fx["b"], fx["L"], fx["U"], fx["var"], solver)
    assert np.allclose(Ln, fx["ld"]("signature_embeddings_updated").T)
    Un = corrnmf.update_sample_embeddings(fx["aux"], fx["a"], fx["b"], fx["L"], fx["U"], fx["var"], solver)
    assert np.allclose(Un, fx["ld"]("sample_embeddings_updated").T)


def test_own_newton_cg_equals_scipy_on_random_problems():
    """600 embedding problems of the CorrNMF kind (aux consistent with the exponential model, like inside a fit)."""
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(300):
        m, n_other = int(rng.integers(1, 7)), int(rng.integers(2, 40))
        others = rng.normal(size=(n_other, m))
        s, s_others = float(rng.normal()), rng.normal(size=n_other)
        var = float(rng.uniform(0.3, 3.0))
        e_true = rng.normal(size=m) * 0.7
        aux_vec = rng.poisson(np.exp(s + s_others + others @ e_true) * 20 + 1.0).astype(float) / 20
        e0 = rng.normal(size=m)
        for maxiter in (3, None):
            a = corrnmf.update_embedding(e0, others, s, s_others, var, aux_vec, maxiter, "scipy")
            b = corrnmf.update_embedding(e0, others, s, s_others, var, aux_vec, maxiter, "own")
            worst = max(worst, float(np.max(np.abs(a - b) / (np.abs(a) + 1e-6))))
    assert worst < 1e-6, worst


# ---- multimodal correlated NMF (reference tests/test_mmcorrnmf.py:132-330, fixtures models/multimodal_corrnmf) -----
PM = "models/multimodal_corrnmf"


@pytest.fixture
def mm():
    from oracle import mmcorrnmf

    ld = lambda name: np.load(golden_path(PM, f"{name}.npy"))  # noqa: E731
    mods = []
    for n in range(2):
        mods.append(
            dict(
                X=load_counts(PM, f"model{n}_counts.csv").T.astype(float),
                W=ld(f"model{n}_signatures_mat_init").T,
                a=ld(f"model{n}_signature_scalings_init"),
                b=ld(f"model{n}_sample_scalings_init"),
                L=ld(f"model{n}_signature_embeddings_init").T,
            )
        )
    U = ld("sample_embeddings_init").T
    mmcorrnmf.compute_exposures(mods, U)
    return dict(mods=mods, U=U, var=float(ld("variance_init")), ld=ld, mm=mmcorrnmf)


def test_mm_elbo_and_aux(mm):
    assert np.allclose(mm["mm"].elbo(mm["mods"], mm["U"], mm["var"]), mm["ld"]("objective_init"))
    auxs = mm["mm"].compute_auxs(mm["mods"])
    for n, (md, aux) in enumerate(zip(mm["mods"], auxs)):
        p = mm["ld"](f"model{n}_p")  # (V, k, D)
        assert np.allclose(aux, np.einsum("vd,vkd->kd", md["X"].T, p))


@pytest.mark.parametrize("solver", ["scipy", "own"])
def test_mm_updates(mm, solver):
    mods, U, var, ld, M = mm["mods"], mm["U"], mm["var"], mm["ld"], mm["mm"]
    auxs = [np.einsum("vd,vkd->kd", md["X"].T, ld(f"model{n}_p")) for n, md in enumerate(mods)]
    for n, (md, aux) in enumerate(zip(mods, auxs)):
        assert np.allclose(corrnmf.update_sample_scalings(md["X"], md["a"], md["L"], U), ld(f"model{n}_sample_scalings_updated"))
        assert np.allclose(corrnmf.update_signature_scalings(aux, md["b"], md["L"], U), ld(f"model{n}_signature_scalings_updated"))
        assert np.allclose(corrnmf.update_signature_embeddings(aux, md["a"], md["b"], md["L"], U, var, solver),
                           ld(f"model{n}_signature_embeddings_updated").T)
        assert np.allclose(klnmf.update_W(md["X"].T, md["W"].T, md["H"].T).T, ld(f"model{n}_signatures_mat_updated").T)
    assert np.allclose(M.update_sample_embeddings(mods, auxs, U, var, solver), ld("sample_embeddings_updated").T)
    assert np.allclose(M.update_variance(mods, U), ld("variance_updated"))


@pytest.mark.parametrize("tag", ["corrnmf_pcawg_k4_dim3_seed3", "corrnmf_pcawg_k6_dim2_seed8"])
def test_oracle_reproduces_live_reference_corrnmf_trajectory(tag):
    """Whole CorrNMFDet iterations of the LIVE reference (oracle/make_golden.py::corrnmf_case, run in the build container):
    our initialisation gives the reference's starting point for the seed, and the oracle's iterations reproduce its ELBO
    history and final parameters."""
    import os

    import pandas as pd
    from conftest import ROOT

    import salamander_b200 as sal
    from oracle import EPSILON
    from salamander_b200.initialization.initialize import initialize_corrnmf

    z = np.load(os.path.join(ROOT, "tests", "golden", "trajectories", f"{tag}.npz"))
    k, dim, seed, n_iter = int(z["k"]), int(z["dim"]), int(z["seed"]), int(z["n_iter"])
    cnt = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T
    adata = sal.AnnData(cnt)
    adata.X = np.asarray(adata.X, dtype=float).clip(EPSILON)
    np.random.seed(seed)
    asig, var0 = initialize_corrnmf(adata, k, dim, "random", None, seed=seed)
    assert np.array_equal(np.asarray(asig.X), z["W0"])
    assert np.array_equal(np.asarray(asig.obsm["embeddings"]), z["L0"]) and np.array_equal(np.asarray(adata.obsm["embeddings"]), z["U0"])
    assert np.array_equal(np.asarray(asig.obs["scalings"].values, dtype=float), z["a0"]) and float(var0) == float(z["var0"])
    X = np.asarray(adata.X, dtype=float)
    W, a, b, L, U, var = z["W0"].copy(), z["a0"].copy(), z["b0"].copy(), z["L0"].copy(), z["U0"].copy(), float(z["var0"])
    hist = []
    for _ in range(n_iter):
        W, a, b, L, U, var, H = corrnmf.update_parameters(X, W, a, b, L, U, var)
        hist.append(corrnmf.elbo(X, W, H, L, U, var))
    assert np.allclose(hist, z["history"], rtol=1e-9, atol=0), (hist, z["history"])
    assert np.allclose(W, z["W"], rtol=1e-7, atol=1e-14) and np.allclose(a, z["a"], rtol=1e-7) and np.allclose(b, z["b"], rtol=1e-7)
    assert np.allclose(L, z["L"], rtol=1e-6, atol=1e-9) and np.allclose(U, z["U"], rtol=1e-6, atol=1e-9)
    assert np.isclose(var, float(z["var"]), rtol=1e-9)


def _mm_frames():
    import os

    import pandas as pd
    from conftest import ROOT

    from oracle import EPSILON

    data = os.path.join(ROOT, "salamander_b200", "data")
    return {name: pd.read_csv(os.path.join(data, f"pcawg_breast_{name}.csv"), index_col=0).T.astype(float).clip(lower=EPSILON)
            for name in ("sbs", "indel", "sv")}


def test_oracle_reproduces_live_reference_mmcorrnmf_trajectory():
    """Whole MultimodalCorrNMF iterations of the LIVE reference on the three PCAWG modalities
    (oracle/make_golden.py::mmcorrnmf_case): our initialisation gives the reference's start for the seed, the oracle's
    iterations its ELBO history and final parameters."""
    import os

    from conftest import ROOT

    import salamander_b200 as sal
    from oracle import mmcorrnmf as mm

    z = np.load(os.path.join(ROOT, "tests", "golden", "trajectories", "mmcorrnmf_pcawg_ns322_dim2_seed5.npz"))
    ns, dim, seed, n_iter = [int(v) for v in z["ns"]], int(z["dim"]), int(z["seed"]), int(z["n_iter"])
    frames = _mm_frames()
    mdata = sal.MuData({name: sal.AnnData(df) for name, df in frames.items()})
    from salamander_b200.initialization.initialize import initialize_mmcorrnmf

    np.random.seed(seed)
    asignatures, var_init = initialize_mmcorrnmf(mdata, ns, dim, "random", None, seed=seed)
    assert np.array_equal(np.asarray(mdata.obsm["embeddings"]), z["U0"]) and float(var_init) == float(z["var0"])
    mods = []
    for name in ("sbs", "indel", "sv"):
        asig = asignatures[name]
        assert np.array_equal(np.asarray(asig.X), z[f"{name}_W0"]), name
        assert np.array_equal(np.asarray(asig.obsm["embeddings"]), z[f"{name}_L0"]), name
        mods.append(dict(X=np.asarray(frames[name].values, dtype=float), W=z[f"{name}_W0"].copy(), a=z[f"{name}_a0"].copy(),
                         b=z[f"{name}_b0"].copy(), L=z[f"{name}_L0"].copy()))
    U, var = z["U0"].copy(), float(z["var0"])
    mm.compute_exposures(mods, U)
    hist = []
    for _ in range(n_iter):
        U, var = mm.update_parameters(mods, U, var)
        hist.append(mm.elbo(mods, U, var))
    assert np.allclose(hist, z["history"], rtol=1e-9, atol=0), (hist, z["history"])
    for md, name in zip(mods, ("sbs", "indel", "sv")):
        assert np.allclose(md["W"], z[f"{name}_W"], rtol=1e-7, atol=1e-14), name
        assert np.allclose(md["L"], z[f"{name}_L"], rtol=1e-6, atol=1e-9), name
        assert np.allclose(md["b"], z[f"{name}_b"], rtol=1e-7), name
    assert np.allclose(U, z["U"], rtol=1e-6, atol=1e-9) and np.isclose(var, float(z["var"]), rtol=1e-9)
