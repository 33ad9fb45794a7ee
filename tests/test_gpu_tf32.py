"""
GPU parity of the tensor-core (tcgen05 kind::tf32) flavour of the fused pass, through the C ABI.

tf32 operands carry 11 significant bits (round to nearest), accumulation is fp32: single updates are
compared with a float64 torch restatement of the same formulas at rtol 4e-3; whole fits are held to
the north-star fp32 criteria (final KL within 1e-4 relative, signature cosine >= 0.9999) against
trajectories of the LIVE reference (tests/golden/trajectories).  The first test reads the kernel's
diagnostic dump so a wrong descriptor shows up as "which GEMM is wrong" instead of a bad number.
"""

import os

import numpy as np
import pandas as pd
import pytest
import torch
from conftest import GOLDEN, ROOT

import salamander_b200 as sal
from salamander_b200 import AnnData
from salamander_b200._device import PASS_OBJECTIVE, PASS_UPDATE_H, PASS_WNUM, Workspace

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float32).eps)
RTOL = 4e-3


def _problem(D, k, seed, dev):
    gen = torch.Generator(device=dev).manual_seed(seed)
    W = torch.rand((k, 96), generator=gen, device=dev, dtype=torch.float64) + 0.01
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=gen, device=dev, dtype=torch.float64) * 400 + 1.0
    X = torch.poisson(H @ W, generator=gen).clamp_min(EPS)
    return X, W, H


def _run(X, W, H, flags, debug=False):
    dev = X.device
    D, k = H.shape
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    Xf, Wf, Hf = X.float().contiguous(), W.float().contiguous(), H.float().contiguous()
    Hout = torch.full_like(Hf, -1.0)
    Wnum = torch.full_like(Wf, -1.0)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    dbg = None
    if debug:
        dbg = torch.zeros(128 * 96 + 128 * 32, dtype=torch.float32, device=dev)
        ws.set_debug_buffer(dbg)
    ws.klnmf_pass(Xf, Wf, Hf, flags, H_out=Hout, Wnum=Wnum, objective=obj)
    torch.cuda.synchronize()
    ws.set_debug_buffer(None)
    ws.close()
    return Hout.double(), Wnum.double(), float(obj.item()), dbg


def _relerr(a, b):
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())


def test_stage_by_stage_first_tile():
    dev = torch.device("cuda:0")
    X, W, H = _problem(300, 20, 1, dev)
    Hout, Wnum, _, dbg = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM, debug=True)
    R_ref = X / (H @ W)
    R = dbg[: 128 * 96].reshape(128, 96).double()
    Hn = dbg[128 * 96 :].reshape(128, 32).double()[:, :20]
    e_r = _relerr(R, R_ref[:128])
    e_hn = _relerr(Hn, R_ref[:128] @ W.T)
    e_h = _relerr(Hout, (H * (R_ref @ W.T)).clamp_min(EPS))
    e_w = _relerr(Wnum, H.T @ R_ref)
    print(f"tf32 stage errors: R {e_r:.2e} (G1 + divide)  Hn {e_hn:.2e} (G2)  H_out {e_h:.2e}  Wnum {e_w:.2e} (G3)")
    assert e_r < RTOL, f"G1 / quotient wrong: {e_r}"
    assert e_hn < RTOL, f"G2 wrong: {e_hn}"
    assert e_h < RTOL, f"H update wrong: {e_h}"
    assert e_w < RTOL, f"G3 / numerator wrong: {e_w}"


@pytest.mark.parametrize(
    "D,k",
    [(1, 4), (127, 8), (128, 12), (129, 16), (1000, 20), (4097, 24), (777, 28), (50_000, 32), (200_003, 20)]
    # k % 4 != 0: exposure tiles go through the 3-D tensor-map view, the partial last tile through plain loads / stores
    + [(1, 1), (100, 2), (128, 3), (129, 5), (1000, 7), (4096, 9), (4097, 13), (20_000, 17), (50_001, 21), (777, 26), (100_003, 30), (256, 31)],
)
def test_pass_matches_float64(D, k):
    dev = torch.device("cuda:0")
    X, W, H = _problem(D, k, 100 + k, dev)
    R = X / (H @ W)
    kl_ref = float((X * torch.log(R) - X + H @ W).sum())
    Hout, Wnum, obj, _ = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM | PASS_OBJECTIVE)
    assert _relerr(Hout, (H * (R @ W.T)).clamp_min(EPS)) < RTOL
    assert _relerr(Wnum, H.T @ R) < RTOL
    assert abs(obj - kl_ref) / abs(kl_ref) < RTOL
    # each flag alone
    Hout2, _, _, _ = _run(X, W, H, PASS_UPDATE_H)
    assert torch.equal(Hout2, Hout)
    _, Wnum2, _, _ = _run(X, W, H, PASS_WNUM)
    assert torch.equal(Wnum2, Wnum)
    # objective-only passes: same error-compensated WH; the logarithm is lg2.approx (2^-22), whose error is a smooth
    # function of the iterate (harmless for the convergence test) but, because sum x ln r and sum (wh - x) cancel to
    # ~1 % of their size at a random start, shows up as a few 1e-6 relative on the objective
    _, _, obj2, _ = _run(X, W, H, PASS_OBJECTIVE)
    assert abs(obj2 - kl_ref) / abs(kl_ref) < 2e-5
    assert abs(obj2 - obj) / abs(obj) < 1e-9


def test_deterministic_and_in_place():
    dev = torch.device("cuda:0")
    X, W, H = _problem(70_001, 20, 7, dev)
    a = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM)
    b = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # H_out aliasing H_in (how the models call it)
    ws = Workspace(96, 70_001, 20, torch.float32, dev, math="tf32_always")
    Hf = H.float().contiguous()
    Wnum = torch.empty((20, 96), dtype=torch.float32, device=dev)
    ws.klnmf_pass(X.float().contiguous(), W.float().contiguous(), Hf, PASS_UPDATE_H | PASS_WNUM, H_out=Hf, Wnum=Wnum)
    torch.cuda.synchronize()
    assert torch.equal(Hf.double(), a[0]) and torch.equal(Wnum.double(), a[1])
    ws.close()


def test_generic_k_deterministic_and_in_place():
    """k % 4 != 0 with a partial last tile: repeatable bit for bit, and H_out may alias H_in."""
    dev = torch.device("cuda:0")
    D, k = 70_001, 11
    X, W, H = _problem(D, k, 9, dev)
    a = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM)
    b = _run(X, W, H, PASS_UPDATE_H | PASS_WNUM)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    Hf = H.float().contiguous()
    Wnum = torch.empty((k, 96), dtype=torch.float32, device=dev)
    ws.klnmf_pass(X.float().contiguous(), W.float().contiguous(), Hf, PASS_UPDATE_H | PASS_WNUM, H_out=Hf, Wnum=Wnum)
    torch.cuda.synchronize()
    assert torch.equal(Hf.double(), a[0]) and torch.equal(Wnum.double(), a[1])
    ws.close()


def test_unsupported_shapes_use_the_exact_kernels():
    """V != 96 (and per-sample KL, Poisson likelihood ...): the call still succeeds through the exact FMA kernels -- never a
    CPU fallback."""
    dev = torch.device("cuda:0")
    D, k, V = 500, 8, 83
    gen = torch.Generator(device=dev).manual_seed(3)
    W = torch.rand((k, V), generator=gen, device=dev, dtype=torch.float64) + 0.01
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=gen, device=dev, dtype=torch.float64) * 400 + 1.0
    X = torch.poisson(H @ W, generator=gen).clamp_min(EPS)
    ws = Workspace(V, D, k, torch.float32, dev, math="tf32_always")
    Hout = torch.empty((D, k), dtype=torch.float32, device=dev)
    Wnum = torch.empty((k, V), dtype=torch.float32, device=dev)
    ws.klnmf_pass(X.float().contiguous(), W.float().contiguous(), H.float().contiguous(), PASS_UPDATE_H | PASS_WNUM, H_out=Hout, Wnum=Wnum)
    torch.cuda.synchronize()
    ws.close()
    R = X / (H @ W)
    assert _relerr(Hout.double(), (H * (R @ W.T)).clamp_min(EPS)) < 3e-5
    assert _relerr(Wnum.double(), H.T @ R) < 3e-5


def _pcawg():
    counts = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0)
    return AnnData(counts.T)


@pytest.mark.parametrize("tag", ["klnmf_pcawg_k8_seed5", "klnmf_pcawg_k4_seed6"])
def test_fit_meets_fp32_criteria(tag):
    z = np.load(os.path.join(GOLDEN, "trajectories", f"{tag}.npz"))
    kw = {}
    for name in ("min_iterations", "max_iterations"):
        if f"ctor_{name}" in z.files:
            kw[name] = int(z[f"ctor_{name}"])
    model = sal.models.KLNMF(n_signatures=int(z["k"]), init_method="random", dtype="float32", math="tf32", **kw)
    model.fit(_pcawg(), init_kwargs={"seed": int(z["seed"])})
    hist, ref = np.array(model.history["objective_function"]), z["history"]
    A, B = model.asignatures.X, z["W"]
    cos = np.sum(A * B, axis=1) / (np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1))
    print(f"{tag}: final KL {hist[-1]:.6f} vs reference {ref[-1]:.6f} (rel {abs(hist[-1] - ref[-1]) / ref[-1]:.2e}), "
          f"{len(hist)} vs {len(ref)} checkpoints, min cosine {cos.min():.7f}")
    assert abs(hist[-1] - ref[-1]) / abs(ref[-1]) < 1e-4
    # k = 8 on 192 samples has a nearly flat direction (8,740 reference iterations, objective steps of ~1e-6 relative per test):
    # fp32 iterates stop a little earlier on the same plateau and still meet 0.9999 -- provided the objective is not evaluated
    # in single precision (tests/test_fp32_plateau.py shows both for the reference arithmetic; csrc/klnmf_small.cu)
    assert cos.min() >= 0.9999


def test_large_fit_tf32_matches_float64_fit():
    """At a sample count where the tensor-core path is actually taken (D >= SAL_TF32_MIN_SAMPLES) a whole fit
    in tf32 mode is compared with the float64 fit (itself pinned to the live reference at 1e-9) from the same
    start: same stopping behaviour, final KL within 1e-4 relative, signature cosine >= 0.9999."""
    import bench

    D, k = 20_000, 8
    X = bench.synth_rows(0, D, k).astype(np.float64)
    W0, H0 = bench.init_rows(X, 0, k)
    fits = {}
    for dtype, math in (("float64", "fma"), ("float32", "tf32")):
        model = sal.models.KLNMF(n_signatures=k, init_method="custom", min_iterations=500, max_iterations=4000, dtype=dtype, math=math)
        model.fit(AnnData(X.copy()), init_kwargs={"signatures_mat": W0.copy(), "exposures_mat": H0.copy()})
        fits[dtype] = model
    a, b = fits["float64"], fits["float32"]
    ha, hb = np.array(a.history["objective_function"]), np.array(b.history["objective_function"])
    A, B = a.asignatures.X, b.asignatures.X
    cos = np.sum(A * B, axis=1) / (np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1))
    n = min(len(ha), len(hb))
    print(f"fp64: {a.n_iterations} iterations, KL {ha[-1]:.4f}; tf32: {b.n_iterations} iterations, KL {hb[-1]:.4f}; "
          f"max rel gap over the common history {np.max(np.abs(ha[:n] - hb[:n]) / ha[:n]):.2e}; min cosine {cos.min():.7f}")
    assert np.max(np.abs(ha[:n] - hb[:n]) / ha[:n]) < 1e-4
    assert abs(ha[-1] - hb[-1]) / ha[-1] < 1e-4
    assert cos.min() >= 0.9999


@pytest.mark.parametrize("D,k", [(5000, 8), (70_001, 20), (33_333, 7)])
def test_mvnmf_pass_variants(D, k):
    """The two passes only MvNMF issues: WNUM | HSUM | OBJECTIVE (numerator, row sums of H, previous objective) and the
    line-search trial OBJECTIVE | UPDATE_H with h_scale (H is read as clip(H * scale), written back like that)."""
    from salamander_b200._device import PASS_HSUM

    dev = torch.device("cuda:0")
    X, W, H = _problem(D, k, 50 + k, dev)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    Xf, Wf, Hf = X.float().contiguous(), W.float().contiguous(), H.float().contiguous()
    Wnum = torch.empty((k, 96), dtype=torch.float32, device=dev)
    hsum = torch.empty(k, dtype=torch.float32, device=dev)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    n0 = ws.launches
    ws.klnmf_pass(Xf, Wf, Hf, PASS_WNUM | PASS_HSUM | PASS_OBJECTIVE, Wnum=Wnum, hsum=hsum, objective=obj)
    torch.cuda.synchronize()
    assert ws.launches - n0 == 3  # tensor-core pass, row-sum partials, reduction -- not the FMA kernel (2 launches)
    R = X / (H @ W)
    kl_ref = float((X * torch.log(R) - X + H @ W).sum())
    assert _relerr(Wnum.double(), H.T @ R) < RTOL
    assert _relerr(hsum.double(), H.sum(0)) < 1e-5
    assert abs(float(obj.item()) - kl_ref) / abs(kl_ref) < RTOL

    scale = torch.rand(k, dtype=torch.float64, device=dev) + 0.5
    Hs = (H * scale).clamp_min(EPS)
    Hout = torch.full_like(Hf, -1.0)
    ws.klnmf_pass(Xf, Wf, Hf, PASS_OBJECTIVE | PASS_UPDATE_H, H_out=Hout, h_scale=scale.float(), objective=obj)
    torch.cuda.synchronize()
    Rs = X / (Hs @ W)
    kl_s = float((X * torch.log(Rs) - X + Hs @ W).sum())
    assert _relerr(Hout.double(), Hs) < 1e-6
    assert abs(float(obj.item()) - kl_s) / abs(kl_s) < 2e-5
    # the same trial with the NEXT H step fused in (SAL_PASS_SCALED_UPDATE): objective of the rescaled exposures, output = their
    # multiplicative update
    from salamander_b200._device import PASS_SCALED_UPDATE

    Hnext = torch.full_like(Hf, -1.0)
    ws.klnmf_pass(Xf, Wf, Hf, PASS_OBJECTIVE | PASS_UPDATE_H | PASS_SCALED_UPDATE, H_out=Hnext, h_scale=scale.float(), objective=obj)
    torch.cuda.synchronize()
    assert abs(float(obj.item()) - kl_s) / abs(kl_s) < 2e-5
    assert _relerr(Hnext.double(), (Hs * (Rs @ W.T)).clamp_min(EPS)) < RTOL
    # ... and on the exact kernels in float64: bit-identical to the two separate passes
    ws64 = Workspace(96, D, k, torch.float64, dev)
    Hs64, Hn_two, Hn_one = torch.empty_like(H), torch.empty_like(H), torch.empty_like(H)
    o1, o2 = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
    ws64.klnmf_pass(X, W, H, PASS_OBJECTIVE | PASS_UPDATE_H, H_out=Hs64, h_scale=scale, objective=o1)
    ws64.klnmf_pass(X, W, Hs64, PASS_UPDATE_H, H_out=Hn_two)
    ws64.klnmf_pass(X, W, H, PASS_OBJECTIVE | PASS_UPDATE_H | PASS_SCALED_UPDATE, H_out=Hn_one, h_scale=scale, objective=o2)
    torch.cuda.synchronize()
    assert torch.equal(Hn_one, Hn_two) and float(o1.item()) == float(o2.item())
    ws64.close()
    ws.close()


def test_mvnmf_fit_on_tensor_cores_tracks_float64():
    """MvNMF with math='tf32' (all three passes of an iteration on the tensor-core kernel) against the float64 fit from
    the same start: final penalised objective within 1e-4, signatures cosine >= 0.9999."""
    import bench

    X = bench.synth_rows(0, 20_000, 8).astype(np.float64)
    out = {}
    for name, kw in {"f64": dict(dtype="float64"), "tf32": dict(dtype="float32", math="tf32")}.items():
        m = sal.models.MvNMF(n_signatures=8, init_method="random", lam=1.0, delta=1.0, min_iterations=150, max_iterations=150, **kw)
        m.fit(AnnData(X.copy()), init_kwargs={"seed": 3})
        out[name] = (m.history["objective_function"][-1], np.array(m.asignatures.X))
    (o64, W64), (o32, W32) = out["f64"], out["tf32"]
    assert abs(o32 - o64) / abs(o64) < 1e-4, (o32, o64)
    cos = np.sum(W64 * W32, axis=1) / (np.linalg.norm(W64, axis=1) * np.linalg.norm(W32, axis=1))
    assert cos.min() >= 0.9999, cos


@pytest.mark.parametrize("use_kl,use_lhalf", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("D,k", [(9000, 8), (40_001, 13)])
def test_weighted_pass_on_tensor_cores(D, k, use_kl, use_lhalf):
    """weights_kl / weights_lhalf (reference _utils_klnmf.py:333-360, klnmf.py:75-79) through the tensor-core kernel: weighted
    numerator, weighted objective + l-half term, closed-form H update."""
    dev = torch.device("cuda:0")
    X, W, H = _problem(D, k, 70 + k, dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    w = torch.rand(D, generator=gen, device=dev, dtype=torch.float64) + 0.5 if use_kl else None
    lam = torch.rand(D, generator=gen, device=dev, dtype=torch.float64) * 3.0 if use_lhalf else None
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    Hout = torch.full((D, k), -1.0, dtype=torch.float32, device=dev)
    Wnum = torch.empty((k, 96), dtype=torch.float32, device=dev)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    n0 = ws.launches
    ws.klnmf_pass(X.float().contiguous(), W.float().contiguous(), H.float().contiguous(), PASS_UPDATE_H | PASS_WNUM | PASS_OBJECTIVE,
                  H_out=Hout, Wnum=Wnum, objective=obj, w_kl=None if w is None else w.float(), w_lhalf=None if lam is None else lam.float())
    torch.cuda.synchronize()
    assert ws.launches - n0 == 2
    ws.close()
    A = X / (H @ W)
    wd = torch.ones(D, dtype=torch.float64, device=dev) if w is None else w
    Hn = A @ W.T
    if lam is None:
        H_ref = (H * Hn).clamp_min(EPS)
    else:
        wsq = (wd * wd)[:, None]
        root = 0.5 * lam[:, None] - torch.sqrt(0.25 * lam[:, None] ** 2 + 4.0 * H * Hn * wsq)
        H_ref = (0.25 * root * root / wsq).clamp_min(EPS)
    kl_rows = (X * torch.log(A) - X + H @ W).sum(1)
    obj_ref = float((wd * kl_rows).sum()) + (0.0 if lam is None else float((lam[:, None] * torch.sqrt(H)).sum()))
    assert _relerr(Wnum.double(), (H * wd[:, None]).T @ A) < RTOL
    assert _relerr(Hout.double(), H_ref) < (2e-2 if use_lhalf else RTOL)  # the closed form amplifies tf32 noise where root ~ 0
    assert abs(float(obj.item()) - obj_ref) / abs(obj_ref) < RTOL
