"""
GPU parity of the persistent period kernel (sal_klnmf_period: a whole convergence-test period of joint updates in one
cooperative launch, reference signature_nmf.py:361-380 around update_WH / kl_divergence, _utils_klnmf.py:11-55, 281-361).

* one launch of L updates == L launches of the two-kernel update (sal_klnmf_update) on the same tensor-core arithmetic:
  only the summation order of the per-CTA numerator partials differs (fixed in both), so W / H agree to fp32 rounding
  and the fused objectives to double rounding;
* against the float64 oracle restatement (oracle/klnmf.py) within the tf32 tolerance;
* the multi-GPU exchange protocol with two EMULATED ranks inside one cooperative launch on one GPU
  (sal_klnmf_period_emulated; kernels that wait on one another must never be separate launches on one device):
  bit-identical W on both ranks, equal to the single-rank result on the concatenated shard to rounding, >= 100
  consecutive updates with tags / slots rolling over many times.
"""

import numpy as np
import pytest
import torch

from salamander_b200._device import Workspace, klnmf_period_emulated

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float32).eps)


def _problem(D, k, seed, dev):
    gen = torch.Generator(device=dev).manual_seed(seed)
    W = torch.rand((k, 96), generator=gen, device=dev, dtype=torch.float64) + 0.01
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=gen, device=dev, dtype=torch.float64) * 400 + 1.0
    X = torch.poisson(H @ W, generator=gen).clamp_min(EPS)
    return X.float().contiguous(), W.float().contiguous(), H.float().contiguous()


def _relerr(a, b):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())


def _two_kernel_updates(ws, X, W, H, L, n_given, every):
    """L updates through sal_klnmf_update (pass + reduction kernel), objective of every `every`-th incoming iterate."""
    W, H = W.clone(), H.clone()
    W2, Wnum = torch.empty_like(W), torch.empty_like(W)
    objs = []
    for u in range(L):
        obj = None
        if every and u % every == 0:
            obj = torch.zeros(1, dtype=torch.float64, device=X.device)
        ws.klnmf_update(X, W, W2, H, H, n_given, True, Wnum, objective=obj)
        W, W2 = W2, W
        if obj is not None:
            objs.append(obj)
    torch.cuda.synchronize()
    return W, H, [float(o.item()) for o in objs]


@pytest.mark.parametrize(
    "D,k,L,n_given",
    [
        (20000, 20, 10, 0),   # bench shape, 157 tiles over 148 CTAs
        (4099, 5, 7, 0),      # generic k (3-D exposure view), ragged last tile, fewer tiles than SMs
        (12345, 8, 10, 1),    # one given signature
        (9000, 30, 4, 0),     # generic k, two X stages
        (50000, 32, 3, 0),    # largest k
        (300, 12, 5, 0),      # 3 tiles: three CTAs, slices of 384 values reduced in chunks
        (100, 16, 5, 2),      # a single CTA
    ],
)
def test_period_equals_two_kernel_updates(D, k, L, n_given):
    dev = torch.device("cuda", 0)
    X, W, H = _problem(D, k, 11 + D + k, dev)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    assert ws.period_supported(n_given, 1)
    every = 3
    W_ref, H_ref, obj_ref = _two_kernel_updates(ws, X, W, H, L, n_given, every)
    final = torch.zeros(1, dtype=torch.float64, device=dev)
    from salamander_b200._device import PASS_OBJECTIVE

    ws.klnmf_pass(X, W_ref, H_ref, PASS_OBJECTIVE, objective=final)
    n_obj = -(-L // every) + 1
    W_out, H_out = torch.full_like(W, -1.0), torch.full_like(H, -1.0)
    objs = torch.full((n_obj,), -1.0, dtype=torch.float64, device=dev)
    ws.klnmf_period(X, W, W_out, H, H_out, n_given, True, L, every, True, objectives=objs)
    torch.cuda.synchronize()
    assert _relerr(W_out, W_ref) < 2e-6, _relerr(W_out, W_ref)
    assert _relerr(H_out, H_ref) < 2e-5, _relerr(H_out, H_ref)
    got = objs.cpu().numpy()
    assert np.isclose(got[0], obj_ref[0], rtol=1e-7, atol=0), (got, obj_ref)  # same inputs; the fp32 per-row sums associate differently
    assert np.allclose(got[:-1], obj_ref, rtol=1e-6, atol=0), (got, obj_ref)
    assert np.isclose(got[-1], float(final.item()), rtol=1e-6), (got[-1], float(final.item()))
    if n_given:
        assert torch.equal(W_out[:n_given], W[:n_given].clamp_min(EPS))
    # deterministic, and in place (H_out == H_in, W_out == W_in) gives the same bits
    W2, H2 = W.clone(), H.clone()
    objs2 = torch.zeros_like(objs)
    ws.klnmf_period(X, W2, W2, H2, H2, n_given, True, L, every, True, objectives=objs2)
    torch.cuda.synchronize()
    assert torch.equal(W2, W_out) and torch.equal(H2, H_out) and torch.equal(objs2, objs)
    ws.close()


def test_period_against_the_float64_oracle():
    """Ten joint updates from a random start against oracle.klnmf.update_WH in float64 (tf32 tolerance) and the fused
    objective against oracle.klnmf.kl_divergence."""
    from oracle import klnmf as oracle_klnmf

    dev = torch.device("cuda", 0)
    D, k, L = 6000, 10, 10
    X, W, H = _problem(D, k, 5, dev)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    W_out, H_out = torch.empty_like(W), torch.empty_like(H)
    objs = torch.zeros(3, dtype=torch.float64, device=dev)
    ws.klnmf_period(X, W, W_out, H, H_out, 0, True, L, 5, True, objectives=objs)
    torch.cuda.synchronize()
    Xn, Wn, Hn = X.double().cpu().numpy().T, W.double().cpu().numpy().T, H.double().cpu().numpy().T
    kls = []
    for u in range(L):
        if u % 5 == 0:
            kls.append(oracle_klnmf.kl_divergence(Xn, Wn, Hn))
        Wn, Hn = oracle_klnmf.update_WH(Xn, Wn, Hn)
    kls.append(oracle_klnmf.kl_divergence(Xn, Wn, Hn))
    assert np.allclose(objs.cpu().numpy(), kls, rtol=1e-4), (objs.cpu().numpy(), kls)
    assert np.allclose(W_out.double().cpu().numpy().T, Wn, rtol=2e-2, atol=1e-7)
    cos = (W_out.double().cpu().numpy() * Wn.T).sum(1) / np.linalg.norm(W_out.double().cpu().numpy(), axis=1) / np.linalg.norm(Wn.T, axis=1)
    assert cos.min() > 0.99999, cos
    ws.close()


def _emulated(X, W, H, split, L, every, n_launches, dev):
    D, k = H.shape
    shards = [(0, split), (split, D)]
    wss = [Workspace(96, hi - lo, k, torch.float32, dev, math="tf32_always") for lo, hi in shards]
    nbytes = int(wss[0].lib.sal_p2p_exchange_bytes(32, 2))
    recv = [torch.zeros(nbytes // 4, dtype=torch.int32, device=dev) for _ in shards]
    table = torch.tensor([r.data_ptr() for r in recv], dtype=torch.int64, device=dev)
    states = [torch.tensor([1, 0], dtype=torch.int32, device=dev) for _ in shards]
    Xs = [X[lo:hi].contiguous() for lo, hi in shards]
    Ws = [W.clone() for _ in shards]
    Hs = [H[lo:hi].clone() for lo, hi in shards]  # (a leading slice is a view: the updates are in place)
    n_obj = -(-L // every) + 1
    objs = [torch.zeros(n_obj, dtype=torch.float64, device=dev) for _ in shards]
    all_objs = []
    for _ in range(n_launches):
        klnmf_period_emulated(wss, Xs, Ws, Ws, Hs, Hs, 0, True, L, every, True, objs, [table, table], states)
        torch.cuda.synchronize()
        assert torch.equal(objs[0], objs[1])
        all_objs.append(objs[0].clone())
    assert int(states[0][0]) == 1 + n_launches * (L + 1) and int(states[1][0]) == int(states[0][0])
    for w in wss:
        w.close()
    return Ws, torch.cat(Hs), torch.stack(all_objs)


def test_two_emulated_ranks_exchange_inside_the_kernel():
    dev = torch.device("cuda", 0)
    D, k, L, every, n_launches = 16001, 20, 10, 5, 10  # 100 updates + 10 objective sweeps: 110 tags, slots roll over 55 times
    X, W, H = _problem(D, k, 3, dev)
    Ws, Hcat, objs = _emulated(X, W, H, 7000, L, every, n_launches, dev)
    assert torch.equal(Ws[0], Ws[1]), "replicas must stay bit-identical"
    # the same updates on one rank holding everything
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    W1, H1 = W.clone(), H.clone()
    o1 = torch.zeros(objs.shape[1], dtype=torch.float64, device=dev)
    ref_objs = []
    for _ in range(n_launches):
        ws.klnmf_period(X, W1, W1, H1, H1, 0, True, L, every, True, objectives=o1)
        torch.cuda.synchronize()
        ref_objs.append(o1.clone())
    ws.close()
    assert np.allclose(objs.cpu().numpy(), torch.stack(ref_objs).cpu().numpy(), rtol=1e-6)
    # 100 updates amplify the last-bit differences of the two summation orders: tolerances are those of an fp32 iteration
    assert _relerr(Ws[0], W1) < 1e-3, _relerr(Ws[0], W1)
    cos = torch.nn.functional.cosine_similarity(Ws[0].double(), W1.double(), dim=1)
    assert float(cos.min()) > 1 - 1e-9
    assert float(((Hcat.double() - H1.double()).abs().sum() / H1.double().abs().sum())) < 1e-4
    # deterministic
    Ws_b, Hcat_b, objs_b = _emulated(X, W, H, 7000, L, every, n_launches, dev)
    assert torch.equal(Ws_b[0], Ws[0]) and torch.equal(Hcat_b, Hcat) and torch.equal(objs_b, objs)


def test_two_emulated_ranks_one_update_is_exact():
    """After ONE update the only difference to a single rank is the association of the numerator sum."""
    dev = torch.device("cuda", 0)
    D, k = 9000, 7
    X, W, H = _problem(D, k, 8, dev)
    Ws, Hcat, objs = _emulated(X, W, H, 4500, 1, 1, 1, dev)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    W1, H1 = torch.empty_like(W), torch.empty_like(H)
    o1 = torch.zeros(2, dtype=torch.float64, device=dev)
    ws.klnmf_period(X, W, W1, H, H1, 0, True, 1, 1, True, objectives=o1)
    torch.cuda.synchronize()
    ws.close()
    assert torch.equal(Ws[0], Ws[1])
    assert _relerr(Ws[0], W1) < 1e-6
    assert torch.equal(Hcat, H1)  # the exposure update only depends on the (identical) incoming W
    got, ref = objs[0].cpu().numpy(), o1.cpu().numpy()
    assert np.isclose(got[0], ref[0], rtol=1e-12)  # incoming iterate: same per-sample terms, another association
    assert np.isclose(got[1], ref[1], rtol=1e-6)   # after the update: W differs in the last bits


def test_headline_config_against_the_oracle():
    """BASELINE config 3 at full size: 96 x 1,000,000, k = 20, dtype float32 / math tf32, 20 iterations of KLNMF.fit from
    bench.init_rows against oracle.klnmf_mt.HostKLNMF (float64, same arithmetic as oracle.klnmf.update_WH / kl_divergence,
    reference _utils_klnmf.py:11-55, 281-361) on the FULL matrix.  North-star fp32 criteria: final KL within 1e-4 relative,
    per-signature cosine >= 0.9999."""
    import bench
    import salamander_b200 as sal
    from oracle.klnmf_mt import HostKLNMF
    from salamander_b200 import AnnData

    D, k, n_iter = 1_000_000, 20, 20
    X32 = bench.synth_rows(0, D, k)
    W0, H0 = bench.init_rows(X32, 0, k)
    model = sal.models.KLNMF(n_signatures=k, init_method="custom", min_iterations=n_iter, max_iterations=n_iter, dtype="float32", math="tf32")
    model.fit(AnnData(X32), init_kwargs={"signatures_mat": W0.copy(), "exposures_mat": H0.copy()})
    assert model.launch_stats["driver"] == "persistent period kernel"
    kl_gpu = model.history["objective_function"][-1]
    host = HostKLNMF(X32.astype(np.float64))
    W, H = W0.copy(), H0.copy()
    for _ in range(n_iter):
        W, H = host.update_WH(W, H)
    kl_cpu = host.kl_divergence(W, H)
    host.close()
    A = np.asarray(model.asignatures.X)
    cos = (A * W).sum(1) / (np.linalg.norm(A, axis=1) * np.linalg.norm(W, axis=1))
    print(f"96 x 1M, k = 20, {n_iter} iterations: KL gpu {kl_gpu:.3f} cpu {kl_cpu:.3f} (rel {abs(kl_gpu - kl_cpu) / kl_cpu:.2e}), min cosine {cos.min():.9f}")
    assert abs(kl_gpu - kl_cpu) / abs(kl_cpu) < 1e-4
    assert cos.min() >= 0.9999
    Hg = np.asarray(model.adata.obsm["exposures"])
    assert np.abs(Hg - H).sum() / np.abs(H).sum() < 1e-3


@pytest.mark.parametrize("D,k,n_given", [(4099, 5, 0), (12345, 20, 1), (300, 12, 0)])
def test_no_writes_outside_the_outputs(D, k, n_given):
    """Guard bands (compute-sanitizer is not available on the GPU pool): every output buffer of the period kernel and of the
    two-kernel update lies between sentinel regions that must come back untouched; inputs must not change."""
    dev = torch.device("cuda", 0)
    X, W, H = _problem(D, k, 21 + D, dev)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    pad = 4096  # floats on either side (16-byte alignment of the views is kept)
    SENT = -12345.0

    def banded(n):
        buf = torch.full((n + 2 * pad,), SENT, dtype=torch.float32, device=dev)
        return buf, buf[pad : pad + n]

    Wb, W_out = banded(k * 96)
    Hb, H_out = banded(D * k)
    ob = torch.full((4 + 2 * 8,), SENT, dtype=torch.float64, device=dev)
    objs = ob[8:12]
    X0, W0, H0 = X.clone(), W.clone(), H.clone()
    ws.klnmf_period(X, W, W_out.view(k, 96), H, H_out.view(D, k), n_given, True, 6, 2, True, objectives=objs)
    torch.cuda.synchronize()
    for buf, n in ((Wb, k * 96), (Hb, D * k)):
        assert bool((buf[:pad] == SENT).all()) and bool((buf[pad + n :] == SENT).all())
    assert bool((ob[:8] == SENT).all()) and bool((ob[12:] == SENT).all())
    assert bool((objs > 0).all()) and bool(torch.isfinite(W_out).all()) and bool(torch.isfinite(H_out).all())
    assert torch.equal(X, X0) and torch.equal(W, W0) and torch.equal(H, H0)
    # the two-kernel update through the same bands
    Wb.fill_(SENT), Hb.fill_(SENT)
    Wnum = torch.empty_like(W)
    ws.klnmf_update(X, W, W_out.view(k, 96), H, H_out.view(D, k), n_given, True, Wnum)
    torch.cuda.synchronize()
    for buf, n in ((Wb, k * 96), (Hb, D * k)):
        assert bool((buf[:pad] == SENT).all()) and bool((buf[pad + n :] == SENT).all())
    assert torch.equal(X, X0) and torch.equal(W, W0) and torch.equal(H, H0)
    ws.close()
