"""
GPU parity at the model level (public API: constructor + fit(adata) + single updates).

* mirrors reference tests/test_klnmf.py:58-91 and tests/test_mvnmf.py:57-90 on the
  reference's golden fixtures;
* full trajectories from the LIVE reference (tests/golden/trajectories, made by
  oracle/make_golden.py): float64 objective history within 1e-9 relative; float32 final KL
  within 1e-4 relative and signature cosine >= 0.9999 (the north-star tolerances).
"""

import os
import pickle

import numpy as np
import pandas as pd
import pytest
from conftest import GOLDEN, ROOT, golden_path

import salamander_b200 as sal
from salamander_b200 import AnnData

pytestmark = pytest.mark.gpu

TRAJ = os.path.join(GOLDEN, "trajectories")


def counts_adata(*parts):
    counts = pd.read_csv(golden_path(*parts), index_col=0)
    return AnnData(counts.T)


def pcawg_adata():
    counts = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0)
    return AnnData(counts.T)


@pytest.fixture(params=[1, 2])
def k(request):
    return request.param


@pytest.fixture(params=["float64", "float32"])
def dtype(request):
    return request.param


def _model_init(cls, path, k, dtype, **kw):
    adata = counts_adata(path, "counts.csv")
    W_init = np.load(golden_path(path, f"W_init_nsigs{k}.npy"))
    H_init = np.load(golden_path(path, f"H_init_nsigs{k}.npy"))
    model = cls(n_signatures=k, dtype=dtype, **kw)
    model.adata = adata
    asig = AnnData(W_init.T)
    asig.var_names = adata.var_names
    model.asignatures = asig
    model.adata.obsm["exposures"] = H_init.T
    return model


# ---- KLNMF (reference tests/test_klnmf.py) -------------------------------------------------


def test_klnmf_objective_function(k, dtype):
    model = _model_init(sal.models.KLNMF, "models/klnmf", k, dtype)
    model._setup_fitting_parameters(None)
    assert np.allclose(model.objective_function(), np.load(golden_path("models/klnmf", f"objective_init_nsigs{k}.npy")))


def test_klnmf_update_parameters(k, dtype):
    model = _model_init(sal.models.KLNMF, "models/klnmf", k, dtype)
    model._setup_fitting_parameters(None)
    model._update_parameters()
    with open(golden_path("models/klnmf", f"WH_updated_joint_nsigs{k}.pkl"), "rb") as f:
        W_updated, H_updated = pickle.load(f)
    assert np.allclose(model.asignatures.X, W_updated.T)
    assert np.allclose(model.adata.obsm["exposures"], H_updated.T)


@pytest.mark.parametrize("cls", ["KLNMF", "MvNMF"])
def test_given_signatures(k, dtype, cls):
    adata = counts_adata("models/klnmf", "counts.csv")
    for n_given in range(1, k + 1):
        given = adata[:n_given, :].copy()
        given.X = given.X / np.sum(given.X, axis=1, keepdims=True)
        model = getattr(sal.models, cls)(n_signatures=k, min_iterations=3, max_iterations=3, dtype=dtype)
        model.fit(adata.copy(), given_parameters={"asignatures": given})
        assert np.allclose(given.X, model.asignatures.X[:n_given, :])
        assert list(model.asignatures.obs_names[:n_given]) == list(given.obs_names)


def test_fit_errors():
    adata = counts_adata("models/klnmf", "counts.csv")
    with pytest.raises(ValueError):
        sal.models.KLNMF(init_method="bogus")
    with pytest.raises(TypeError):
        sal.models.KLNMF().fit(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        sal.models.KLNMF(n_signatures=2).fit(adata.copy(), fitting_kwargs={"weights": 1.0})
    with pytest.raises(ValueError):
        sal.models.KLNMF(n_signatures=2).fit(adata.copy(), fitting_kwargs={"weights_kl": -np.ones(adata.n_obs)})
    with pytest.raises(ValueError):
        sal.models.KLNMF(n_signatures=2).fit(adata.copy(), given_parameters={"signatures": 1})


# ---- MvNMF (reference tests/test_mvnmf.py) -------------------------------------------------


def test_mvnmf_objective_function(k, dtype):
    model = _model_init(sal.models.MvNMF, "models/mvnmf", k, dtype)
    assert np.allclose(model.objective_function(), np.load(golden_path("models/mvnmf", f"objective_init_nsigs{k}.npy")))


def test_mvnmf_update_W(k, dtype):
    model = _model_init(sal.models.MvNMF, "models/mvnmf", k, dtype)
    model._gamma = 1.0
    model._update_W()
    assert np.allclose(model.asignatures.X, np.load(golden_path("models/mvnmf", f"W_updated_nsigs{k}.npy")).T, rtol=1e-5 if dtype == "float64" else 1e-4)


def test_mvnmf_update_H(k, dtype):
    model = _model_init(sal.models.MvNMF, "models/mvnmf", k, dtype)
    model._update_H()
    assert np.allclose(model.adata.obsm["exposures"], np.load(golden_path("models/mvnmf", f"H_updated_nsigs{k}.npy")).T)


# ---- trajectories of the live reference ----------------------------------------------------


def _ctor(z, name, default):
    key = f"ctor_{name}"
    return z[key].item() if key in z.files else default


def _cosine(A, B):
    return np.sum(A * B, axis=1) / (np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1))


@pytest.mark.parametrize("tag", ["klnmf_pcawg_k5_seed0", "klnmf_pcawg_k4_weights", "klnmf_pcawg_k6_given2"])
def test_klnmf_trajectory(tag, dtype):
    z = np.load(os.path.join(TRAJ, f"{tag}.npz"))
    adata = pcawg_adata()
    kk, n_given = int(z["k"]), int(z["n_given"])
    given = None
    if n_given:
        g = pcawg_adata()[:n_given, :].copy()
        g.X = g.X / g.X.sum(axis=1, keepdims=True)
        given = {"asignatures": g}
    fk = {}
    if z["weights_kl"].size:
        fk["weights_kl"] = z["weights_kl"]
    if z["weights_lhalf"].size:
        fk["weights_lhalf"] = z["weights_lhalf"]
    model = sal.models.KLNMF(
        n_signatures=kk,
        init_method="random",
        min_iterations=_ctor(z, "min_iterations", 500),
        max_iterations=_ctor(z, "max_iterations", 10000),
        dtype=dtype,
    )
    model.fit(adata, given_parameters=given, init_kwargs={"seed": int(z["seed"])}, fitting_kwargs=fk or None)
    hist = np.array(model.history["objective_function"])
    ref = z["history"]
    if dtype == "float64":
        assert len(hist) == len(ref)
        assert np.allclose(hist, ref, rtol=1e-9, atol=0)
        assert np.allclose(model.asignatures.X, z["W"], rtol=1e-6, atol=1e-12)
        assert np.allclose(model.adata.obsm["exposures"], z["H"], rtol=1e-6, atol=1e-9)
    else:
        assert abs(hist[-1] - ref[-1]) / abs(ref[-1]) < 1e-4
        assert _cosine(model.asignatures.X, z["W"]).min() >= 0.9999
    # the API contract of the reference: results live in the AnnData objects as float64 host arrays
    assert model.asignatures.X.dtype == np.float64 and model.asignatures.X.shape == (kk, 96)
    assert model.adata.obsm["exposures"].shape == (192, kk)
    assert model.signatures.shape == (96, kk) and model.exposures.shape == (192, kk)


@pytest.mark.parametrize("driver", ["single-CTA persistent kernel", "run-ahead line search", "host loop"])
@pytest.mark.parametrize("tag", ["mvnmf_pcawg_k10_seed0", "mvnmf_pcawg_k3_lam50"])
def test_mvnmf_trajectory(tag, dtype, driver):
    """Every fit driver of MvNMF against the live-reference trajectory: the persistent single-CTA kernel (small problems), the
    run-ahead driver of the larger ones (optimistic line search + roll-back when the full step is rejected) and the plain
    reference loop with a host decision per trial."""
    z = np.load(os.path.join(TRAJ, f"{tag}.npz"))
    model = sal.models.MvNMF(
        n_signatures=int(z["k"]),
        init_method="random",
        lam=_ctor(z, "lam", 1.0),
        delta=_ctor(z, "delta", 1.0),
        min_iterations=_ctor(z, "min_iterations", 500),
        max_iterations=_ctor(z, "max_iterations", 10000),
        dtype=dtype,
    )
    model.use_small_kernel = driver == "single-CTA persistent kernel"
    model.run_ahead = driver == "run-ahead line search"
    model.fit(pcawg_adata(), init_kwargs={"seed": int(z["seed"])})
    if driver != "host loop":
        assert model.launch_stats["driver"] == driver, model.launch_stats
    hist = np.array(model.history["objective_function"])
    ref = z["history"]
    if dtype == "float64":
        assert np.allclose(hist, ref, rtol=1e-9, atol=0)
        assert np.isclose(model._gamma, float(z["gamma"]))
        assert np.allclose(model.asignatures.X, z["W"], rtol=1e-6, atol=1e-12)
    else:
        assert abs(hist[-1] - ref[-1]) / abs(ref[-1]) < 1e-4
        assert _cosine(model.asignatures.X, z["W"]).min() >= 0.9999


@pytest.mark.parametrize("k,lam,delta,seed", [(3, 1e4, 0.1, 3), (6, 1e4, 1e-3, 0)])
def test_mvnmf_run_ahead_rolls_back_rejected_steps(k, lam, delta, seed):
    """A strongly penalised problem in which about every second iteration rejects the full step: the run-ahead driver has to
    drop what it queued on top of the optimistic iterate, restore the buffers and back-track like the reference
    (mvnmf.py:69-92).  Objective history, gamma and the final factors against the oracle."""
    from oracle import EPSILON
    from oracle import mvnmf as omv
    from salamander_b200.initialization.initialize import initialize_mat

    adata = pcawg_adata()
    Xp = np.asarray(adata.X, dtype=float).clip(EPSILON)
    W0, H0 = initialize_mat(Xp, k, "random", seed=seed)
    W, H, gamma, n, hist = omv.fit_mvnmf(Xp.T, W0.T, H0.T, lam=lam, delta=delta, min_iterations=60, max_iterations=60)
    assert gamma < 1.0  # the oracle did back-track
    model = sal.models.MvNMF(n_signatures=k, init_method="random", lam=lam, delta=delta, min_iterations=60, max_iterations=60, dtype="float64")
    model.use_small_kernel = False
    model.fit(adata, init_kwargs={"seed": seed})
    assert model.launch_stats["driver"] == "run-ahead line search" and model.launch_stats["back_tracked_iterations"] >= 10
    assert np.allclose(model.history["objective_function"], hist, rtol=1e-9, atol=0)
    assert np.isclose(model._gamma, gamma, rtol=1e-12)
    assert np.allclose(model.asignatures.X, W.T, rtol=1e-6, atol=1e-12)
    assert np.allclose(model.adata.obsm["exposures"], H.T, rtol=1e-6, atol=1e-10)


def test_reconstruction_error(dtype):
    from oracle import EPSILON, klnmf

    adata = pcawg_adata()
    model = sal.models.KLNMF(n_signatures=3, init_method="random", min_iterations=20, max_iterations=20, dtype=dtype)
    model.fit(adata, init_kwargs={"seed": 4})
    ref = klnmf.samplewise_kl_divergence(adata.X.T, model.asignatures.X.T, adata.obsm["exposures"].T)
    model.compute_reconstruction_errors()
    assert np.allclose(adata.obs["reconstruction_error"].values, ref, rtol=1e-9 if dtype == "float64" else 2e-3, atol=1e-2 if dtype == "float32" else 0)
    assert np.isclose(model.reconstruction_error, ref.sum(), rtol=1e-5)


def test_custom_init_rescaled_on_device_equals_host_path(monkeypatch):
    """Large custom exposure matrices are normalised / clipped on the device (sal_scale_clip_rows) instead of on the
    host (reference initialize.py:116-118); both routes must give the same fit."""
    from salamander_b200.initialization import initialize

    rng = np.random.default_rng(11)
    adata = pcawg_adata()
    W0 = rng.dirichlet(np.ones(96), size=4) * rng.uniform(0.5, 2.0, size=(4, 1))  # rows deliberately not normalised
    H0 = rng.gamma(1.0, 50.0, size=(192, 4))
    H0[3, 2] = 0.0  # clipped to EPSILON by either route
    res = []
    for min_size in (1 << 22, 1):
        monkeypatch.setattr(initialize, "DEFER_MIN_SIZE", min_size)
        model = sal.models.KLNMF(n_signatures=4, init_method="custom", min_iterations=30, max_iterations=30, dtype="float64")
        model.fit(adata.copy(), init_kwargs={"signatures_mat": W0.copy(), "exposures_mat": H0.copy()})
        res.append((model.asignatures.X, model.adata.obsm["exposures"], model.history["objective_function"]))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-13) and np.allclose(res[0][1], res[1][1], rtol=1e-13)
    assert np.allclose(res[0][2], res[1][2], rtol=1e-13)


def test_sweep_over_k_and_restarts():
    """Model-selection sweep (tutorial.ipynb:1975-2013 semantics): every (k, seed) job equals the stand-alone fit."""
    from salamander_b200.sweep import error_curve, sweep_klnmf

    adata = pcawg_adata()
    table, best = sweep_klnmf(adata, [2, 3], n_restarts=2, seed0=5, min_iterations=40, max_iterations=40, dtype="float64")
    assert list(table["n_signatures"]) == [2, 2, 3, 3] and list(table["seed"]) == [5, 6, 5, 6]
    solo = sal.models.KLNMF(n_signatures=3, init_method="random", min_iterations=40, max_iterations=40, dtype="float64")
    solo.fit(pcawg_adata(), init_kwargs={"seed": 6})
    row = table[(table.n_signatures == 3) & (table.seed == 6)].iloc[0]
    assert np.isclose(row.reconstruction_error, solo.reconstruction_error, rtol=1e-12)
    curve = error_curve(table)
    assert curve.loc[3] < curve.loc[2]
    assert set(best) == {2, 3} and np.isclose(best[3].reconstruction_error, curve.loc[3])


def test_resident_sweep_equals_standalone_fits():
    """A sweep keeps ONE clipped device copy of the counts for all of its fits, draws the random exposures on the device from
    the resident totals and downloads only the fits that are the best of their k: every row of the table must equal the
    stand-alone fit with the same (k, seed), and best[k] must carry that fit's results."""
    import bench
    from salamander_b200.sweep import sweep_klnmf

    X = bench.synth_rows(0, 6000, 6)
    X[0, :5] = 0.0  # an entry below EPSILON: the sweep rebinds adata.X to the clipped matrix once
    adata = AnnData(X.copy())
    kw = dict(min_iterations=30, max_iterations=30, dtype="float32", math="tf32", init_device=True)
    table, best = sweep_klnmf(adata, [3, 4], n_restarts=2, seed0=1, **kw)
    assert float(np.asarray(adata.X).min()) > 0
    assert len(table) == 4 and set(best) == {3, 4}
    for _, row in table.iterrows():
        m = sal.models.KLNMF(n_signatures=int(row.n_signatures), init_method="random", replica=True, **kw)
        m.errors_in_fit = True
        m.fit(AnnData(X.copy()), init_kwargs={"seed": int(row.seed)})
        assert np.isclose(m.history["objective_function"][-1], row.objective, rtol=1e-6)
        assert np.isclose(m.reconstruction_error, row.reconstruction_error, rtol=1e-6)
        b = best[int(row.n_signatures)]
        if np.isclose(b.reconstruction_error, row.reconstruction_error, rtol=1e-12):
            assert np.allclose(b.asignatures.X, m.asignatures.X, rtol=1e-4, atol=1e-9)
            assert np.allclose(b.adata.obsm["exposures"], m.adata.obsm["exposures"], rtol=1e-3, atol=1e-5)
