"""
GPU parity of MultimodalCorrNMF through the public API, mirroring reference tests/test_mmcorrnmf.py:132-347 on the
reference's golden fixtures (tests/golden/models/multimodal_corrnmf: 2 modalities, ns_signatures [2, 3], dim 2),
plus whole iterations against the oracle on the three PCAWG modalities (SBS-96, indel-83, SV-32).
"""

import os

import numpy as np
import pandas as pd
import pytest
from conftest import ROOT, golden_path

import salamander_b200 as sal
from salamander_b200 import AnnData, MuData

from oracle import mmcorrnmf as oracle

pytestmark = pytest.mark.gpu
P = "models/multimodal_corrnmf"
NS, DIM = [2, 3], 2


def ld(name):
    return np.load(golden_path(P, f"{name}.npy"))


def make_mdata():
    adatas = {f"mod{n}": AnnData(pd.read_csv(golden_path(P, f"model{n}_counts.csv"), index_col=0).T) for n in range(2)}
    mdata = MuData(adatas)
    mdata.obsm["embeddings"] = ld("sample_embeddings_init").T
    for n in range(2):
        mdata[f"mod{n}"].obs["scalings"] = ld(f"model{n}_sample_scalings_init")
    return mdata


@pytest.fixture
def model_init():
    mdata = make_mdata()
    asignatures = {}
    for n in range(2):
        asigs = AnnData(ld(f"model{n}_signatures_mat_init").T)
        asigs.var_names = mdata[f"mod{n}"].var_names
        asigs.obs["scalings"] = ld(f"model{n}_signature_scalings_init")
        asigs.obsm["embeddings"] = ld(f"model{n}_signature_embeddings_init").T
        asignatures[f"mod{n}"] = asigs
    model = sal.models.MultimodalCorrNMF(ns_signatures=NS, dim_embeddings=DIM)
    model.mdata = mdata
    model.asignatures = asignatures
    model.compute_exposures()
    model.variance = float(ld("variance_init"))
    return model


@pytest.fixture
def auxs():
    out = {}
    for n in range(2):
        X = pd.read_csv(golden_path(P, f"model{n}_counts.csv"), index_col=0).values  # (V, D)
        out[f"mod{n}"] = np.einsum("vd,vkd->kd", X, ld(f"model{n}_p"))
    return out


def test_init_signature_names(model_init):
    given = {}
    for mod_name, adata in model_init.mdata.mod.items():
        asigs = AnnData(np.zeros((1, adata.n_vars)))
        asigs.obs_names = ["A"]
        asigs.var_names = adata.var_names
        given[mod_name] = {"asignatures": asigs}
    model_init._initialize(given)
    for mod_name, asigs in model_init.asignatures.items():
        for k, sig_name in enumerate(asigs.obs_names):
            assert sig_name == ("A" if k == 0 else f"{mod_name} Sig{k}")


def test_objective_function(model_init):
    assert np.allclose(model_init.objective_function(), ld("objective_init"))


def test_compute_auxs(model_init, auxs):
    got = model_init._compute_auxs()
    for name in auxs:
        assert np.allclose(got[name], auxs[name])


def test_modality_updates(model_init, auxs):
    model_init.update_signatures()
    for n in range(2):
        assert np.allclose(model_init.asignatures[f"mod{n}"].X, ld(f"model{n}_signatures_mat_updated").T)


def test_update_sample_scalings(model_init):
    model_init.update_sample_scalings()
    for n in range(2):
        assert np.allclose(model_init.mdata[f"mod{n}"].obs["scalings"].values, ld(f"model{n}_sample_scalings_updated"))


def test_update_signature_scalings(model_init, auxs):
    model_init.update_signature_scalings(auxs)
    for n in range(2):
        assert np.allclose(model_init.asignatures[f"mod{n}"].obs["scalings"].values, ld(f"model{n}_signature_scalings_updated"))


def test_update_signature_embeddings(model_init, auxs):
    model_init.update_signature_embeddings(auxs)
    for n in range(2):
        assert np.allclose(model_init.asignatures[f"mod{n}"].obsm["embeddings"], ld(f"model{n}_signature_embeddings_updated").T)


def test_update_sample_embeddings(model_init, auxs):
    model_init.update_sample_embeddings(auxs)
    assert np.allclose(model_init.mdata.obsm["embeddings"], ld("sample_embeddings_updated").T)


def test_update_variance(model_init):
    model_init.update_variance()
    assert np.allclose(model_init.variance, ld("variance_updated"))


@pytest.mark.parametrize("ns,dim", [([1, 2], 1), ([2, 2], 2)])
def test_given_parameters(ns, dim):
    rng = np.random.default_rng(1)

    def model():
        return sal.models.MultimodalCorrNMF(ns_signatures=ns, dim_embeddings=dim, max_iterations=3, init_method="random")

    mdata = make_mdata()
    m0, m1 = mdata.mod.keys()
    for n_given in range(1, ns[0] + 1):
        given = mdata.mod[m0][:n_given, :].copy()
        given.X = given.X.astype(float)
        given.X /= np.sum(given.X, axis=1, keepdims=True)
        gp = {m0: {"asignatures": given}}
        mod = model().fit(make_mdata(), given_parameters=gp, init_kwargs={"seed": 2})
        assert np.allclose(given.X, mod.asignatures[m0].X[:n_given, :])
        assert not np.allclose(given.X, mod.asignatures[m1].X[:n_given, :])
        if n_given < ns[0]:
            other = mod.asignatures[m0].X[n_given:, :].copy()
            mod._update_parameters(gp)
            assert not np.allclose(other, mod.asignatures[m0].X[n_given:, :])
    v = rng.uniform(size=ns[0])
    mod = model().fit(make_mdata(), given_parameters={m0: {"signature_scalings": v}}, init_kwargs={"seed": 2})
    assert np.allclose(v, mod.asignatures[m0].obs["scalings"])
    v = rng.uniform(size=(ns[0], dim))
    mod = model().fit(make_mdata(), given_parameters={m0: {"signature_embeddings": v}}, init_kwargs={"seed": 2})
    assert np.allclose(v, mod.asignatures[m0].obsm["embeddings"])
    v = rng.uniform(size=mdata.n_obs)
    mod = model().fit(make_mdata(), given_parameters={m0: {"sample_scalings": v}}, init_kwargs={"seed": 2})
    assert np.allclose(v, mod.mdata.mod[m0].obs["scalings"]) and not np.allclose(v, mod.mdata.mod[m1].obs["scalings"])
    v = rng.uniform(size=(mdata.n_obs, dim))
    mod = model().fit(make_mdata(), given_parameters={"sample_embeddings": v}, init_kwargs={"seed": 2})
    assert np.allclose(v, mod.mdata.obsm["embeddings"])
    mod = model().fit(make_mdata(), given_parameters={"variance": 3.0}, init_kwargs={"seed": 2})
    assert np.allclose(3.0, mod.variance)
    with pytest.raises(KeyError):
        model().fit(make_mdata(), given_parameters={m0: {"variance": 2.0}})
    with pytest.raises(ValueError):
        model().fit(make_mdata(), given_parameters={"bogus": 1})
    with pytest.raises(TypeError):
        model().fit(np.zeros((3, 3)))


def test_iterations_match_the_oracle_on_pcawg_modalities():
    data = os.path.join(ROOT, "salamander_b200", "data")
    # MultimodalCorrNMF does not clip the counts (reference mmcorrnmf.py:196-209); some samples have no SV at all, so the
    # counts are clipped here the way SignatureNMF._setup_adata would (zeros -> EPSILON) to keep every log finite
    frames = {
        name: pd.read_csv(os.path.join(data, f"pcawg_breast_{name}.csv"), index_col=0).T.astype(float).clip(lower=1.1920928955078125e-07)
        for name in ("sbs", "indel", "sv")
    }
    ns, dim, n_iter = [3, 2, 2], 2, 4
    mdata = MuData({name: AnnData(df) for name, df in frames.items()})
    model = sal.models.MultimodalCorrNMF(ns_signatures=ns, dim_embeddings=dim, init_method="random")
    model._setup_mdata(mdata)
    np.random.seed(5)
    model._initialize(None, {"seed": 5})
    mods = []
    for name in model.mod_names:
        ad_, as_ = mdata[name], model.asignatures[name]
        mods.append(dict(X=np.asarray(ad_.X, dtype=float), W=np.array(as_.X), a=np.array(as_.obs["scalings"].values, dtype=float),
                         b=np.array(ad_.obs["scalings"].values, dtype=float), L=np.array(as_.obsm["embeddings"])))
    U, var = np.array(mdata.obsm["embeddings"]), float(model.variance)
    oracle.compute_exposures(mods, U)
    hist_ref = []
    for _ in range(n_iter):
        U, var = oracle.update_parameters(mods, U, var)
        hist_ref.append(oracle.elbo(mods, U, var))
    with model._resident():
        model._in_fit = True
        hist = []
        for _ in range(n_iter):
            model._update_parameters(None)
            hist.append(model.objective_function())
        model._in_fit = False
    assert np.allclose(hist, hist_ref, rtol=1e-8), (hist, hist_ref)
    for md, name in zip(mods, model.mod_names):
        assert np.allclose(model.asignatures[name].X, md["W"], rtol=1e-6, atol=1e-12)
        assert np.allclose(model.asignatures[name].obsm["embeddings"], md["L"], rtol=1e-5, atol=1e-8)
        assert np.allclose(model.mdata[name].obs["scalings"].values, md["b"], rtol=1e-6)
    assert np.allclose(model.mdata.obsm["embeddings"], U, rtol=1e-5, atol=1e-7)
    assert np.isclose(model.variance, var, rtol=1e-7)


def test_fit_matches_the_live_reference_trajectory():
    """``MultimodalCorrNMF.fit`` against whole iterations of the LIVE reference on the three PCAWG modalities
    (tests/golden/trajectories/mmcorrnmf_pcawg_ns322_dim2_seed5.npz, oracle/make_golden.py::mmcorrnmf_case)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "trajectories", "mmcorrnmf_pcawg_ns322_dim2_seed5.npz"))
    ns, dim, seed, n_iter = [int(v) for v in z["ns"]], int(z["dim"]), int(z["seed"]), int(z["n_iter"])
    data = os.path.join(ROOT, "salamander_b200", "data")
    frames = {
        name: pd.read_csv(os.path.join(data, f"pcawg_breast_{name}.csv"), index_col=0).T.astype(float).clip(lower=1.1920928955078125e-07)
        for name in ("sbs", "indel", "sv")
    }
    mdata = MuData({name: AnnData(df) for name, df in frames.items()})
    model = sal.models.MultimodalCorrNMF(ns_signatures=ns, dim_embeddings=dim, init_method="random", min_iterations=n_iter,
                                         max_iterations=n_iter, conv_test_freq=1)
    np.random.seed(seed)
    model.fit(mdata, init_kwargs={"seed": seed})
    hist = np.array(model.history["objective_function"])
    assert hist.shape == z["history"].shape
    assert np.allclose(hist, z["history"], rtol=1e-8, atol=0), (hist, z["history"])
    for name in ("sbs", "indel", "sv"):
        assert np.allclose(model.asignatures[name].X, z[f"{name}_W"], rtol=1e-6, atol=1e-12), name
        assert np.allclose(model.asignatures[name].obsm["embeddings"], z[f"{name}_L"], rtol=1e-5, atol=1e-8), name
        assert np.allclose(model.mdata[name].obs["scalings"].values, z[f"{name}_b"], rtol=1e-6), name
    assert np.allclose(model.mdata.obsm["embeddings"], z["U"], rtol=1e-5, atol=1e-7)
    assert np.isclose(model.variance, float(z["var"]), rtol=1e-7)
