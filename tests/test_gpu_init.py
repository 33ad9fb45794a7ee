"""NNDSVD initialisation with the SVD on the device (initialization/device_nndsvd.py) against the reference's route,
scikit-learn's ``_initialize_nmf`` (reference initialization/methods.py:69-86)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch
from conftest import ROOT

import salamander_b200 as sal
from salamander_b200 import AnnData
from salamander_b200.initialization.device_nndsvd import init_nndsvd_device
from salamander_b200.initialization.methods import init_nndsvd

pytestmark = pytest.mark.gpu


def _pcawg():
    return pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T


@pytest.mark.parametrize("method", ["nndsvd", "nndsvda", "nndsvdar"])
@pytest.mark.parametrize("k,tol", [(5, 1e-10), (10, 1e-4)])  # k = 10: the randomized SVD's own error is ~1e-5 there
def test_device_nndsvd_matches_scikit_learn(method, k, tol):
    X = _pcawg().values.astype(float)
    s_ref, e_ref = init_nndsvd(X, k, method=method, seed=4)
    after_ref = np.random.random()
    s_dev, e_dev = init_nndsvd_device(X, k, method=method, seed=4, device=torch.device("cuda:0"))
    after_dev = np.random.random()
    assert after_ref == after_dev  # the global numpy RNG is left where scikit-learn leaves it
    assert s_dev.shape == (k, 96) and e_dev.shape == (192, k)
    assert np.array_equal(s_ref == 0, s_dev == 0) and np.array_equal(e_ref == 0, e_dev == 0)
    assert np.abs(s_dev - s_ref).max() <= tol * np.abs(s_ref).max()
    assert np.abs(e_dev - e_ref).max() <= tol * np.abs(e_ref).max()


def test_large_matrix_and_fit_from_device_init():
    """100k synthetic samples: device NNDSVD equals the host route to ~1e-8, and a fit started from it reproduces the
    fit started from the host initialisation (same stopping iteration, objective to 1e-9)."""
    import bench

    X = bench.synth_rows(0, 100_000, 8).astype(np.float64)
    s_ref, e_ref = init_nndsvd(X, 8, seed=0)
    s_dev, e_dev = init_nndsvd_device(X, 8, seed=0, device=torch.device("cuda:0"))
    assert np.abs(s_dev - s_ref).max() <= 1e-7 * np.abs(s_ref).max()
    assert np.abs(e_dev - e_ref).max() <= 1e-7 * np.abs(e_ref).max()

    cnt = _pcawg()
    fits = []
    for init_device in (False, True):
        model = sal.models.KLNMF(n_signatures=5, init_method="nndsvd", dtype="float64", init_device=init_device, max_iterations=2000)
        model.fit(AnnData(cnt))
        fits.append(model)
    assert fits[0].n_iterations == fits[1].n_iterations
    h0, h1 = (np.array(m.history["objective_function"]) for m in fits)
    assert np.allclose(h0, h1, rtol=1e-9, atol=0)
    assert np.allclose(fits[0].asignatures.X, fits[1].asignatures.X, rtol=1e-6, atol=1e-12)


def test_fit_from_device_random_initialisation():
    """init_method='random' with init_device=True: exposures drawn on the device (reproducible for a seed), rescaled and
    clipped on the device; the fit behaves like one from the host initialisation (same final objective within 1 %)."""
    import bench

    X = bench.synth_rows(0, 30_000, 6).astype(np.float64)
    finals = {}
    for name, kw in {"host": {}, "dev1": {"init_device": True}, "dev2": {"init_device": True}}.items():
        model = sal.models.KLNMF(n_signatures=6, init_method="random", dtype="float64", min_iterations=300, max_iterations=300, **kw)
        model.fit(AnnData(X.copy()), init_kwargs={"seed": 5})
        hist = model.history["objective_function"]
        assert hist[-1] < hist[0]
        finals[name] = (hist[-1], np.array(model.adata.obsm["exposures"]))
    assert finals["dev1"][0] == finals["dev2"][0] and np.array_equal(finals["dev1"][1], finals["dev2"][1])
    assert abs(finals["dev1"][0] - finals["host"][0]) / finals["host"][0] < 1e-2
