"""
The small-problem kernels on a thread-block cluster (klnmf_cluster_kernel / mvnmf_cluster_kernel: the samples in blocks over 8 or
16 CTAs, partial sums exchanged through distributed shared memory) against the single-CTA kernels they replace from 64 samples
on: ragged sample counts (last CTA short or empty), every k padding class, given signatures, float32 and float64.  Same
arithmetic, different summation grouping: equal to rounding.  The outputs sit between sentinel bands that must come back untouched
(compute-sanitizer is closed on the GPU pool: this is the memcheck stand-in for the new kernels).
"""

import os

import numpy as np
import pytest
import torch

from salamander_b200._device import Workspace

pytestmark = pytest.mark.gpu
BAND = 4096


def _problem(D, k, seed, dtype):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    W = torch.rand((k, 96), generator=g, device=dev, dtype=torch.float64) + 0.01
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=g, device=dev, dtype=torch.float64) * 300 + 1
    X = torch.poisson(H @ W, generator=g).clamp_min(1.1920928955078125e-07)
    return X.to(dtype).contiguous(), W.to(dtype).contiguous(), H.to(dtype).contiguous()


def _banded(n, dtype, dev):
    buf = torch.full((n + 2 * BAND,), -777.0, dtype=dtype, device=dev)
    return buf, buf[BAND : BAND + n]


def _bands_intact(buf, n):
    return bool((buf[:BAND] == -777.0).all()) and bool((buf[BAND + n :] == -777.0).all())


@pytest.fixture
def cluster_env():
    saved = {v: os.environ.get(v) for v in ("SAL_B200_KLNMF_CLUSTER", "SAL_B200_MVNMF_CLUSTER")}
    yield
    for v, val in saved.items():
        if val is None:
            os.environ.pop(v, None)
        else:
            os.environ[v] = val


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("D,k,n_given", [(64, 3, 0), (65, 5, 1), (100, 10, 0), (192, 20, 2), (255, 7, 0), (256, 32, 0)])
def test_klnmf_cluster_equals_single_cta(D, k, n_given, dtype, cluster_env):
    dev = torch.device("cuda", 0)
    X, W, H = _problem(D, k, 3 * D + k, dtype)
    ws = Workspace(96, D, k, dtype, dev)
    assert ws.small_supported()
    res = {}
    for mode in ("0", "8"):
        os.environ["SAL_B200_KLNMF_CLUSTER"] = mode
        wbuf, Wo = _banded(k * 96, dtype, dev)
        hbuf, Ho = _banded(D * k, dtype, dev)
        obj = torch.zeros(1, dtype=torch.float64, device=dev)
        ws.klnmf_small_updates(X, W, Wo.view(k, 96), H, Ho.view(D, k), n_given, 9, objective=obj)
        torch.cuda.synchronize()
        assert _bands_intact(wbuf, k * 96) and _bands_intact(hbuf, D * k), mode
        res[mode] = (Wo.clone().view(k, 96), Ho.clone().view(D, k), float(obj.item()))
    ws.close()
    tol = 1e-11 if dtype == torch.float64 else 2e-4
    assert abs(res["0"][2] - res["8"][2]) <= 1e-12 * abs(res["0"][2]) if dtype == torch.float64 else abs(res["0"][2] - res["8"][2]) <= 1e-6 * abs(res["0"][2])
    assert torch.allclose(res["0"][0], res["8"][0], rtol=tol, atol=1e-12)
    assert torch.allclose(res["0"][1], res["8"][1], rtol=tol, atol=1e-9)
    assert not torch.equal(res["8"][0], W)  # the updates happened
    if n_given:
        assert torch.equal(res["8"][0][:n_given], W[:n_given].clamp_min(1.1920928955078125e-07))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("D,k,lam,delta", [(64, 3, 1.0, 1.0), (65, 5, 50.0, 0.5), (100, 10, 1.0, 1.0), (192, 12, 1e4, 0.1), (255, 16, 1.0, 1.0), (256, 20, 10.0, 1.0)])
def test_mvnmf_cluster_equals_single_cta(D, k, lam, delta, dtype, cluster_env):
    dev = torch.device("cuda", 0)
    X, W, H = _problem(D, k, 7 * D + k, dtype)
    ws = Workspace(96, D, k, dtype, dev)
    assert ws.mvnmf_small_supported()
    res = {}
    for mode in ("0", "8", "16"):
        os.environ["SAL_B200_MVNMF_CLUSTER"] = mode
        wbuf, Wo = _banded(k * 96, dtype, dev)
        hbuf, Ho = _banded(D * k, dtype, dev)
        obj = torch.zeros(1, dtype=torch.float64, device=dev)
        g_in, g_out = torch.ones(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
        ws.mvnmf_small_updates(X, W, Wo.view(k, 96), H, Ho.view(D, k), lam, delta, 0, 6, g_in, g_out, objective=obj)
        torch.cuda.synchronize()
        assert _bands_intact(wbuf, k * 96) and _bands_intact(hbuf, D * k), mode
        res[mode] = (Wo.clone().view(k, 96), Ho.clone().view(D, k), float(obj.item()), float(g_out.item()))
    ws.close()
    tol = 1e-9 if dtype == torch.float64 else 5e-4
    for mode in ("8", "16"):
        assert np.isclose(res["0"][2], res[mode][2], rtol=1e-12 if dtype == torch.float64 else 1e-6), mode
        if dtype == torch.float64:  # (in float32 a line-search decision on the edge may differ between the summation orders)
            assert res["0"][3] == res[mode][3], (mode, res["0"][3], res[mode][3])
            assert torch.allclose(res["0"][0], res[mode][0], rtol=tol, atol=1e-12), mode
            assert torch.allclose(res["0"][1], res[mode][1], rtol=tol, atol=1e-9), mode
        elif res["0"][3] == res[mode][3]:
            assert torch.allclose(res["0"][0], res[mode][0], rtol=tol, atol=1e-7), mode
    assert not torch.equal(res["8"][0], W)
