"""
Pins the oracle (oracle/klnmf.py, oracle/mvnmf.py) against the reference's OWN golden
fixtures (tests/golden/models/* == reference tests/test_data/models/*), mirroring
reference tests/test_utils_klnmf.py, tests/test_klnmf.py and tests/test_mvnmf.py, and --
when /root/reference is mounted -- against the live reference on random inputs.
CPU only.
"""

import pickle

import numpy as np
import pytest
from conftest import golden_path, load_counts

from oracle import EPSILON, klnmf, mvnmf
from oracle import reference_loader as rl

U = "models/utils_klnmf"


@pytest.fixture(params=[1, 2])
def k(request):
    return request.param


@pytest.fixture
def xwh(k):
    X = load_counts(U, "counts.csv")
    W = np.load(golden_path(U, f"W_nsigs{k}.npy"))
    H = np.load(golden_path(U, f"H_nsigs{k}.npy"))
    return X, W, H


def g(name):
    return np.load(golden_path(U, name))


# ---- reference tests/test_utils_klnmf.py:48-196 ---------------------------------------


def test_kl_divergence(xwh, k):
    assert np.allclose(klnmf.kl_divergence(*xwh), g(f"kl_divergence_nsigs{k}.npy"))


def test_kl_divergence_weights(xwh, k):
    w = 2 * np.ones(xwh[0].shape[1])
    assert np.allclose(klnmf.kl_divergence(*xwh, w), 2 * g(f"kl_divergence_nsigs{k}.npy"))


def test_samplewise_kl(xwh, k):
    assert np.allclose(klnmf.samplewise_kl_divergence(*xwh), g(f"samplewise_kl_divergence_nsigs{k}.npy"))


def test_samplewise_kl_weights(xwh, k):
    w = 2 * np.ones(xwh[0].shape[1])
    w[0] = 3
    got = klnmf.samplewise_kl_divergence(*xwh, w)
    ref = g(f"samplewise_kl_divergence_nsigs{k}.npy")
    assert np.allclose(got[0], 3 * ref[0]) and np.allclose(got[1:], 2 * ref[1:])


def test_poisson_llh(xwh, k):
    assert np.allclose(klnmf.poisson_llh(*xwh), g(f"poisson_llh_nsigs{k}.npy"))


def test_update_W(xwh, k):
    ref = g(f"W_updated_standard_nsigs{k}.npy")
    assert np.allclose(klnmf.update_W(*xwh), ref)
    assert np.allclose(klnmf.update_W(*xwh, 2 * np.ones(xwh[0].shape[1])), ref)


def test_update_W_given(xwh, k):
    X, W, H = xwh
    for n_given in range(1, k + 1):
        out = klnmf.update_W(X, W.copy(), H, n_given_signatures=n_given)
        assert np.array_equal(out[:, :n_given], W[:, :n_given])


def test_update_H(xwh, k):
    ref = g(f"H_updated_standard_nsigs{k}.npy")
    D = xwh[0].shape[1]
    assert np.allclose(klnmf.update_H(*xwh), ref)
    assert np.allclose(klnmf.update_H(*xwh, 2 * np.ones(D), np.zeros(D)), ref)


def test_update_WH(xwh, k):
    Wr, Hr = g(f"W_updated_joint_nsigs{k}.npy"), g(f"H_updated_joint_nsigs{k}.npy")
    D = xwh[0].shape[1]
    for args in [(), (2 * np.ones(D),), (2 * np.ones(D), np.zeros(D))]:
        Wn, Hn = klnmf.update_WH(*xwh, *args)
        assert np.allclose(Wn, Wr) and np.allclose(Hn, Hr)


def test_update_WH_given(xwh, k):
    X, W, H = xwh
    for n_given in range(1, k + 1):
        Wn, _ = klnmf.update_WH(X, W.copy(), H, n_given_signatures=n_given)
        assert np.array_equal(Wn[:, :n_given], W[:, :n_given])


# ---- reference tests/test_klnmf.py:58-74 ----------------------------------------------


def test_klnmf_model_fixtures(k):
    P = "models/klnmf"
    X = load_counts(P, "counts.csv")
    W = np.load(golden_path(P, f"W_init_nsigs{k}.npy"))
    H = np.load(golden_path(P, f"H_init_nsigs{k}.npy"))
    assert np.allclose(klnmf.klnmf_objective(X, W, H), np.load(golden_path(P, f"objective_init_nsigs{k}.npy")))
    with open(golden_path(P, f"WH_updated_joint_nsigs{k}.pkl"), "rb") as f:
        Wr, Hr = pickle.load(f)
    Wn, Hn = klnmf.update_WH(X, W, H)
    assert np.allclose(Wn, Wr) and np.allclose(Hn, Hr)


# ---- reference tests/test_mvnmf.py:57-72 ----------------------------------------------


def test_mvnmf_model_fixtures(k):
    P = "models/mvnmf"
    X = load_counts(P, "counts.csv")
    W = np.load(golden_path(P, f"W_init_nsigs{k}.npy"))
    H = np.load(golden_path(P, f"H_init_nsigs{k}.npy"))
    obj = mvnmf.kl_divergence_penalized(X, W, H, 1.0, 1.0)
    assert np.allclose(obj, np.load(golden_path(P, f"objective_init_nsigs{k}.npy")))
    # _update_H
    assert np.allclose(klnmf.update_H(X, W, H), np.load(golden_path(P, f"H_updated_nsigs{k}.npy")))
    # _update_W = unconstrained + line search with gamma = 1 (H not updated first in the fixture)
    Wu = mvnmf.update_W_unconstrained(X, W, H, 1.0, 1.0)
    Wn, _, _ = mvnmf.line_search(X, W, H, 1.0, 1.0, 1.0, Wu)
    assert np.allclose(Wn, np.load(golden_path(P, f"W_updated_nsigs{k}.npy")))


# ---- live reference (build container only) --------------------------------------------

needs_ref = pytest.mark.skipif(not rl.available(), reason="live reference not mounted")


@needs_ref
@pytest.mark.reference
@pytest.mark.parametrize("kk,D,seed", [(3, 17, 0), (7, 64, 1)])
def test_against_live_reference(kk, D, seed):
    ref = rl.load_utils_klnmf()
    rng = np.random.default_rng(seed)
    V = 96
    X = rng.poisson(20.0, size=(V, D)).astype(float)
    X[rng.random((V, D)) < 0.1] = 0.0
    W = rng.dirichlet(np.ones(V), size=kk).T.copy()
    H = rng.gamma(2.0, 50.0, size=(kk, D))
    wk = rng.uniform(0.5, 2.0, D)
    wl = rng.uniform(0.0, 3.0, D)
    assert np.isclose(klnmf.kl_divergence(X, W, H, wk), ref.kl_divergence(X, W, H, wk), rtol=1e-12)
    assert np.allclose(klnmf.samplewise_kl_divergence(X, W, H, wk), ref.samplewise_kl_divergence(X, W, H, wk), rtol=1e-12)
    assert np.isclose(klnmf.poisson_llh(X, W, H), ref.poisson_llh(X, W, H), rtol=1e-12)
    Xc = X.clip(EPSILON)
    for n_given in (0, 1, kk):
        assert np.allclose(klnmf.update_W(Xc, W, H, wk, n_given), ref.update_W(Xc, W.copy(), H.copy(), wk, n_given), rtol=1e-12, atol=0)
        Wn, Hn = klnmf.update_WH(Xc, W, H, wk, wl, n_given)
        Wr, Hr = ref.update_WH(Xc, W.copy(), H.copy(), wk, wl, n_given)
        assert np.allclose(Wn, Wr, rtol=1e-12, atol=0) and np.allclose(Hn, Hr, rtol=1e-11, atol=0)
        Wn, Hn = klnmf.update_WH(Xc, W, H, None, None, n_given)
        Wr, Hr = ref.update_WH(Xc, W.copy(), H.copy(), None, None, n_given)
        assert np.allclose(Wn, Wr, rtol=1e-12, atol=0) and np.allclose(Hn, Hr, rtol=1e-12, atol=0)
    assert np.allclose(klnmf.update_H(Xc, W, H, wk, wl), ref.update_H(Xc, W.copy(), H.copy(), wk, wl), rtol=1e-11, atol=0)
    assert np.allclose(klnmf.update_H(Xc, W, H), ref.update_H(Xc, W.copy(), H.copy()), rtol=1e-12, atol=0)
