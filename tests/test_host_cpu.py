"""
CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), the product path refuses to run without CUDA, the
threaded CPU port used as bench.py's reference arm equals the oracle, the synthetic workload
does not depend on the rank count, and the sample-sharding helpers work under a 2-process gloo
group (the N > 1 path of SURVEY.md 8(e)).
"""

import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
from conftest import ROOT

from oracle import EPSILON, klnmf
from oracle.klnmf_mt import HostKLNMF


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry

    entry.build()
    from salamander_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "salamander_b200.h")).read()
    declared = set(re.findall(r"\b(sal_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sal_version() >= 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    import salamander_b200 as sal
    from salamander_b200._lib import SalamanderB200Error

    adata = sal.AnnData(np.random.default_rng(0).poisson(5.0, size=(12, 96)).astype(float))
    with pytest.raises(SalamanderB200Error):
        sal.models.KLNMF(n_signatures=2, init_method="random").fit(adata)
    with pytest.raises(SalamanderB200Error):
        sal.models.MvNMF(n_signatures=2, init_method="random").fit(adata)


def test_package_does_not_import_oracle():
    code = "import sys, salamander_b200, salamander_b200.models; assert not any(m.split('.')[0] == 'oracle' for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_threaded_port_equals_oracle():
    rng = np.random.default_rng(3)
    V, D, k = 96, 5000, 7
    W = rng.dirichlet(np.ones(V), size=k)  # [k][V]
    H = rng.gamma(1.0, 100.0, size=(D, k))
    X = rng.poisson(H @ W).astype(float).clip(EPSILON)
    host = HostKLNMF(X, n_threads=3, chunk=777)
    Wr, Hr = klnmf.update_WH(X.T, W.T, H.T)
    kl_r = klnmf.kl_divergence(X.T, W.T, H.T)
    assert np.isclose(host.kl_divergence(W, H), kl_r, rtol=1e-12)
    Wn, Hn = host.update_WH(W, H.copy())
    assert np.allclose(Wn, Wr.T, rtol=1e-11) and np.allclose(Hn, Hr.T, rtol=1e-11)
    Wg, _ = host.update_WH(W, H.copy(), n_given_signatures=2)
    assert np.allclose(Wg, klnmf.update_WH(X.T, W.T, H.T, None, None, 2)[0].T, rtol=1e-11)
    host.close()


def test_synthetic_workload_is_rank_independent():
    import bench

    full = bench.synth_rows(0, 130_000, 20)
    assert full.dtype == np.float32 and full.min() >= np.float32(EPSILON)
    lo, hi = bench.shard_bounds(130_000, 3, 1)
    part = bench.synth_rows(lo, hi, 20)
    assert np.array_equal(part, full[lo:hi])
    W0, H0 = bench.init_rows(full, 0, 20)
    W1, H1 = bench.init_rows(part, lo, 20)
    assert np.array_equal(W0, W1) and np.array_equal(H1, H0[lo:hi])
    assert np.allclose(W0.sum(axis=1), 1.0, atol=1e-4)
    # column totals in the range of PCAWG breast SBS burdens
    assert 2000 < np.median(full.sum(axis=1)) < 12000


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SAL_ROOT"])
from salamander_b200 import _dist
dist.init_process_group("gloo")
rank, world = _dist.world()
assert world == 2
n = 11
lo, hi = _dist.shard_bounds(n, world, rank)
full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
got = _dist.gather_rows(full[lo:hi].clone(), n)
assert torch.equal(got, full), (rank, got)
t = torch.full((4,), float(rank + 1))
_dist.allreduce_sum_(t)
assert torch.equal(t, torch.full((4,), 3.0))
arr = np.full((2, 2), float(rank))
out = _dist.broadcast_numpy(arr, torch.device("cpu"))
assert np.array_equal(out, np.zeros((2, 2)))
# sample-sharded W numerator: the sum of the shards' partial numerators equals the unsharded one
rng = np.random.default_rng(0)
X = rng.poisson(20.0, size=(n, 5)).astype(float) + 1e-3
W = rng.dirichlet(np.ones(5), size=2); H = rng.gamma(2.0, 5.0, size=(n, 2))
num = torch.from_numpy(H[lo:hi].T @ (X[lo:hi] / (H[lo:hi] @ W)))
_dist.allreduce_sum_(num)
assert np.allclose(num.numpy(), H.T @ (X / (H @ W)), rtol=1e-12)
# multi-GPU CorrNMF: signatures are dealt to the ranks for the embedding Newton-CG, the rows are exchanged afterwards
k = 5
j0, j1 = _dist.shard_bounds(k, world, rank)
L = torch.full((k, 3), -1.0, dtype=torch.float64)
L[j0:j1] = torch.arange(j0, j1, dtype=torch.float64)[:, None] + 0.125 * torch.arange(3, dtype=torch.float64)[None, :]
_dist.exchange_owned_rows(L, j0, j1)
assert torch.equal(L, torch.arange(k, dtype=torch.float64)[:, None] + 0.125 * torch.arange(3, dtype=torch.float64)[None, :]), (rank, L)
# [sum L^2 (replicated, counted once), sum U^2 over this rank's samples, sum lnGamma over this rank's samples]
norms = torch.tensor([7.0, float(rank + 1), 10.0 * (rank + 1)], dtype=torch.float64)
_dist.allreduce_sum_counting_replicated_once(norms, slice(0, 1))
assert torch.equal(norms, torch.tensor([7.0, 3.0, 30.0], dtype=torch.float64)), (rank, norms)
# path decisions whose kernels wait on the peers are taken collectively: true only if true on EVERY rank
assert _dist.all_ranks_agree(True, torch.device("cpu")) is True
assert _dist.all_ranks_agree(rank == 0, torch.device("cpu")) is False
assert _dist.all_ranks_agree(False, torch.device("cpu")) is False
# launch numbers of the CorrNMF signature solver's exchange: the same on both ranks, and before the 15-bit tag prefix would
# repeat the receive buffers are zeroed behind a barrier and the count starts over
class _FakeExchange:
    buf = torch.ones(8, dtype=torch.int32)
    launch_id = _dist.LAUNCH_ID_LIMIT - 3
ids = [_dist.next_launch_id(_FakeExchange) for _ in range(4)]
assert ids == [_dist.LAUNCH_ID_LIMIT - 2, _dist.LAUNCH_ID_LIMIT - 1, 1, 2], ids
assert int(_FakeExchange.buf.abs().sum()) == 0
# the period driver's launch plan is a pure function of (min, max, freq): both ranks issue the same launches
from salamander_b200.models.klnmf import period_launch_plan
plan = period_launch_plan(min_iterations=35, max_iterations=100, freq=10, per_launch=16)
assert plan == [(0, 4, False), (4, 1, False), (5, 1, False), (6, 1, False), (7, 1, False), (8, 1, False), (9, 1, True)], plan
assert period_launch_plan(20, 20, 10, 16) == [(0, 2, True)]
assert period_launch_plan(500, 10000, 10, 16)[:4] == [(0, 16, False), (16, 16, False), (32, 16, False), (48, 2, False)]
assert period_launch_plan(0, 25, 10, 16) == [(0, 1, False), (1, 1, False), (2, 1, False)]
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharding_helpers_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, SAL_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script),
    ]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_shard_bounds_cover_everything():
    from salamander_b200._dist import shard_bounds

    for n in (0, 1, 7, 8, 1_000_000):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_signature_matching_and_reorder():
    """match_signatures_pair / match_to_catalog / SignatureNMF.reorder / data_reconstructed (reference utils.py:161-192,
    signature_nmf.py:221-235, 387-406): host-side helpers that align restarts and sweep models."""
    import pandas as pd

    from salamander_b200 import AnnData
    from salamander_b200.models import KLNMF
    from salamander_b200.utils import match_signatures_pair, match_to_catalog

    rng = np.random.default_rng(0)
    W = rng.dirichlet(np.ones(96), size=5)
    perm = np.array([3, 0, 4, 1, 2])
    noisy = W[perm] * (1 + 0.01 * rng.standard_normal((5, 96)))
    s1 = pd.DataFrame(W, index=[f"a{i}" for i in range(5)])
    s2 = pd.DataFrame(noisy, index=[f"b{i}" for i in range(5)])
    order = match_signatures_pair(s1, s2)
    assert np.array_equal(perm[order], np.arange(5))
    assert list(match_to_catalog(s2, s1).index) == [f"a{i}" for i in perm]
    with pytest.raises(ValueError):
        match_signatures_pair(s1, s2.iloc[:4])

    model = KLNMF(n_signatures=5)
    H = rng.random((7, 5))
    model.adata = AnnData(pd.DataFrame(H @ noisy))
    model.adata.obsm["exposures"] = H.copy()
    model.asignatures = AnnData(pd.DataFrame(noisy, index=list(s2.index)))
    recon = model.data_reconstructed
    assert np.allclose(recon.values, H @ noisy)
    model.reorder(AnnData(s1))
    assert np.allclose(model.asignatures.X, noisy[order])
    assert np.allclose(model.adata.obsm["exposures"], H[:, order])
    assert list(model.asignatures.obs_names) == list(s2.index)  # names stay in place unless keep_names
    assert np.allclose(model.adata.obsm["exposures"] @ model.asignatures.X, H @ noisy)


def test_gram_nndsvd_arithmetic_on_the_cpu_device():
    """The NNDSVD construction of initialization/device_nndsvd.py, run with torch's CPU device so that the arithmetic is
    checked without a GPU: equal to scikit-learn's route (what the reference calls, initialization/methods.py:69-86)."""
    import pandas as pd
    import torch

    from salamander_b200.initialization.device_nndsvd import init_nndsvd_device
    from salamander_b200.initialization.methods import init_nndsvd

    X = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T.values.astype(float)
    for method in ("nndsvd", "nndsvdar"):
        s_ref, e_ref = init_nndsvd(X, 5, method=method, seed=2)
        s_dev, e_dev = init_nndsvd_device(X, 5, method=method, seed=2, device=torch.device("cpu"))
        assert np.array_equal(s_ref == 0, s_dev == 0) and np.array_equal(e_ref == 0, e_dev == 0)
        assert np.abs(s_dev - s_ref).max() <= 1e-10 * np.abs(s_ref).max()
        assert np.abs(e_dev - e_ref).max() <= 1e-10 * np.abs(e_ref).max()


def test_device_random_initialisation_contract_on_the_cpu_device():
    """init_random_device (opt-in, models' init_device=True): signatures are the reference's numpy draws for the seed, the
    exposures are reproducible, have the samples' totals as row sums and Dirichlet(1_k) shares -- but are not numpy's draws."""
    import torch

    from salamander_b200.initialization.device_nndsvd import init_random_device
    from salamander_b200.initialization.initialize import initialize_mat
    from salamander_b200.initialization.methods import init_random

    X = np.random.default_rng(0).poisson(5.0, size=(20_000, 96)).astype(np.float64)
    s_ref, e_ref = init_random(X, 6, seed=11)
    s1, e1 = init_random_device(X, 6, seed=11, device=torch.device("cpu"))
    s2, e2 = init_random_device(X, 6, seed=11, device=torch.device("cpu"))
    assert np.array_equal(s1, s_ref) and np.array_equal(e1, e2) and not np.allclose(e1, e_ref)
    assert np.allclose(e1.sum(axis=1), X.sum(axis=1), rtol=1e-12)
    share1, share_ref = e1 / e1.sum(1, keepdims=True), e_ref / e_ref.sum(1, keepdims=True)
    assert np.allclose(share1.mean(0), 1 / 6, atol=5e-3) and np.isclose(share1.var(0).mean(), share_ref.var(0).mean(), rtol=0.05)
    defer = {}
    W, H = initialize_mat(X, 6, "random", None, _defer=defer, seed=11, _init_device=torch.device("cpu"))
    assert "exposure_scale" in defer and np.array_equal(H, e1)  # the rescale / clip of H is left to the device
    assert np.allclose(W.sum(axis=1), 1.0, atol=1e-5)


def test_init_device_routing_and_sweep_containers():
    """Host logic around the optional device initialisation and the sweep: which fits ask for it, and that the fits of a
    sweep share the count matrix instead of copying it."""
    from salamander_b200 import AnnData
    from salamander_b200.models import KLNMF
    from salamander_b200.sweep import _shallow_copy

    X = np.random.default_rng(1).poisson(3.0, size=(50, 96)).astype(np.float64)
    ad = AnnData(X)
    for kwargs in (dict(init_method="nndsvd"), dict(init_method="nndsvd", init_device="auto"), dict(init_method="random", init_device="auto"),
                   dict(init_method="flat", init_device=True)):
        m = KLNMF(n_signatures=3, **kwargs)
        m.adata = ad
        assert m._init_device_kwargs() == {}, kwargs  # host initialisation: default, small matrices, other methods
    with pytest.raises(ValueError):
        KLNMF(n_signatures=3, init_device="yes")
    if not torch.cuda.is_available():  # asking for it without a GPU fails loudly, like every device step
        m = KLNMF(n_signatures=3, init_method="nndsvd", init_device=True)
        m.adata = ad
        with pytest.raises(Exception):
            m._init_device_kwargs()
    shallow = _shallow_copy(ad)
    assert shallow is not ad and np.shares_memory(np.asarray(shallow.X), X)
    assert list(shallow.obs_names) == list(ad.obs_names) and list(shallow.var_names) == list(ad.var_names)
    shallow.obsm["exposures"] = np.zeros((50, 3))
    assert "exposures" not in ad.obsm


def test_initialisation_is_bit_identical_to_the_live_reference():
    """tests/golden/init_pcawg.npz: W0 / H0 of the reference's own ``initialize_mat`` (run in the build container by
    oracle/make_golden.py::init_cases) for every method -- ours must be equal bit for bit, including the numba summation
    order of normalize_WH."""
    import pandas as pd

    from salamander_b200.initialization.initialize import initialize_mat

    z = np.load(os.path.join(ROOT, "tests", "golden", "init_pcawg.npz"))
    X = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T.values.astype(float)
    X = X.clip(EPSILON)
    keys = sorted(k[:-2] for k in z.files if k.endswith("_W"))
    assert len(keys) == 12
    for key in keys:
        method, k = key.rsplit("_k", 1)
        seed = int(z[f"{key}_seed"])
        kw = {} if seed < 0 else {"seed": seed}
        W0, H0 = initialize_mat(X.copy(), int(k), method, **kw)
        assert np.array_equal(W0, z[f"{key}_W"]), key
        assert np.array_equal(H0, z[f"{key}_H"]), key


def test_integration_doc_lists_every_abi_symbol():
    """INTEGRATION.md maps each entry point of include/salamander_b200.h to the reference function it replaces."""
    hdr = open(os.path.join(ROOT, "include", "salamander_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    symbols = sorted(set(re.findall(r"\b(sal_[a-z0-9_]+)\s*\(", hdr)))
    assert len(symbols) >= 30
    assert [s for s in symbols if s not in doc] == []


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exits 0 without a GPU and prints exactly one JSON
    line with the contract's keys; `bench.py` without a GPU refuses loudly instead of falling back to the CPU."""
    import json

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--samples", "40000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not torch.cuda.is_available():
        res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert res.returncode != 0 and "no CPU fallback" in (res.stdout + res.stderr)
