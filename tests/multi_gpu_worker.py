"""
Worker of tests/test_gpu_multi.py (launched with torchrun, one rank per GPU): sample-sharded KLNMF / MvNMF fits
must reproduce the single-process trajectories of the live reference (tests/golden/trajectories) -- the only
exchange per iteration is the all-reduce of the 96 x k numerator (+ scalars), SURVEY.md 8(e).
"""
import os
import sys

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData, _dist as sal_dist  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
TRAJ = os.path.join(ROOT, "tests", "golden", "trajectories")


def pcawg():
    counts = pd.read_csv(os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0)
    return AnnData(counts.T)


z = np.load(os.path.join(TRAJ, "klnmf_pcawg_k5_seed0.npz"))
model = sal.models.KLNMF(n_signatures=5, init_method="random", dtype="float64", device=f"cuda:{local}")
model.fit(pcawg(), init_kwargs={"seed": int(z["seed"])})
hist = np.array(model.history["objective_function"])
assert len(hist) == len(z["history"]), (len(hist), len(z["history"]))
assert np.allclose(hist, z["history"], rtol=1e-9, atol=0), np.max(np.abs(hist - z["history"]) / z["history"])
assert np.allclose(model.asignatures.X, z["W"], rtol=1e-6, atol=1e-12)
assert model.adata.obsm["exposures"].shape == (192, 5)
assert np.allclose(model.adata.obsm["exposures"], z["H"], rtol=1e-6, atol=1e-9)

z = np.load(os.path.join(TRAJ, "mvnmf_pcawg_k3_lam50.npz"))
m2 = sal.models.MvNMF(n_signatures=3, init_method="random", lam=50.0, delta=0.5, min_iterations=300, max_iterations=300,
                      dtype="float64", device=f"cuda:{local}")
m2.fit(pcawg(), init_kwargs={"seed": int(z["seed"])})
h2 = np.array(m2.history["objective_function"])
assert np.allclose(h2, z["history"], rtol=1e-9, atol=0), np.max(np.abs(h2 - z["history"]) / z["history"])
assert np.allclose(m2.asignatures.X, z["W"], rtol=1e-6, atol=1e-12)

# CorrNMFDet, samples sharded: 5 whole iterations (k = 4, dim 3) against the oracle from the same start.  The W numerator,
# the scaling sums, |U|^2 and the likelihood are all-reduced; the signature embeddings are solved by signature shares on
# all-gathered per-sample inputs.
from oracle import corrnmf as oracle_corr  # noqa: E402

k, m, n_iter = 4, 3, 5
cnt = pcawg()
mc = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, init_method="random", min_iterations=n_iter, max_iterations=n_iter,
                           conv_test_freq=1, device=f"cuda:{local}")
mc._setup_adata(cnt)
np.random.seed(3)
mc._initialize(None, {"seed": 3})
X = np.asarray(cnt.X, dtype=float)
W = np.array(mc.asignatures.X)
a, b = np.array(mc.asignatures.obs["scalings"].values, dtype=float), np.array(cnt.obs["scalings"].values, dtype=float)
L, U = np.array(mc.asignatures.obsm["embeddings"]), np.array(cnt.obsm["embeddings"])
var = float(mc.variance)
hist_ref = []
for _ in range(n_iter):
    W, a, b, L, U, var, H = oracle_corr.update_parameters(X, W, a, b, L, U, var)
    hist_ref.append(oracle_corr.elbo(X, W, H, L, U, var))
with mc._resident() as st_c:
    assert st_c.world == world and st_c.D == len(range(*sal_dist.shard_bounds(192, world, rank)))
    assert st_c.sig_exchange() is not None  # the signature-embedding solver exchanges its sums inside the kernel (NVLink)
    mc._in_fit = True
    hist_c = []
    for _ in range(n_iter):
        mc._update_parameters(None)
        hist_c.append(mc.objective_function())
    mc._in_fit = False
assert np.allclose(hist_c, hist_ref, rtol=1e-8), (hist_c, hist_ref)
assert np.allclose(mc.asignatures.X, W, rtol=1e-6, atol=1e-12)
assert np.allclose(mc.asignatures.obs["scalings"].values, a, rtol=1e-6)
assert mc.adata.obs["scalings"].values.shape == (192,)
assert np.allclose(mc.adata.obs["scalings"].values, b, rtol=1e-6)
assert np.allclose(mc.asignatures.obsm["embeddings"], L, rtol=1e-5, atol=1e-8)
assert mc.adata.obsm["embeddings"].shape == (192, m)
assert np.allclose(mc.adata.obsm["embeddings"], U, rtol=1e-5, atol=1e-7)
assert np.isclose(mc.variance, var, rtol=1e-7)
Lc = torch.as_tensor(mc.asignatures.obsm["embeddings"]).cuda()
Lref = Lc.clone()
dist.broadcast(Lref, src=0)
assert torch.equal(Lc, Lref)  # replicated parameters are bit-identical on all ranks

# MultimodalCorrNMF, samples sharded (every modality alike, shared sample embeddings local): whole fit against the
# live-reference trajectory of the three PCAWG modalities
from salamander_b200 import MuData  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "trajectories", "mmcorrnmf_pcawg_ns322_dim2_seed5.npz"))
ns, dim, seed, n_it = [int(v) for v in z["ns"]], int(z["dim"]), int(z["seed"]), int(z["n_iter"])
data_dir = os.path.join(ROOT, "salamander_b200", "data")
frames = {name: pd.read_csv(os.path.join(data_dir, f"pcawg_breast_{name}.csv"), index_col=0).T.astype(float).clip(lower=1.1920928955078125e-07)
          for name in ("sbs", "indel", "sv")}
mm = sal.models.MultimodalCorrNMF(ns_signatures=ns, dim_embeddings=dim, init_method="random", min_iterations=n_it, max_iterations=n_it,
                                  conv_test_freq=1, device=f"cuda:{local}")
np.random.seed(seed)
mm.fit(MuData({name: AnnData(df) for name, df in frames.items()}), init_kwargs={"seed": seed})
assert mm.shard_info["world"] == world and mm.shard_info["local_samples"] < mm.shard_info["samples"], mm.shard_info
assert np.allclose(mm.history["objective_function"], z["history"], rtol=1e-8, atol=0), (mm.history["objective_function"], z["history"])
for name in ("sbs", "indel", "sv"):
    assert np.allclose(mm.asignatures[name].X, z[f"{name}_W"], rtol=1e-6, atol=1e-12), name
    assert np.allclose(mm.asignatures[name].obsm["embeddings"], z[f"{name}_L"], rtol=1e-5, atol=1e-8), name
    assert np.allclose(mm.mdata[name].obs["scalings"].values, z[f"{name}_b"], rtol=1e-6), name
assert np.allclose(mm.mdata.obsm["embeddings"], z["U"], rtol=1e-5, atol=1e-7)
assert np.isclose(mm.variance, float(z["var"]), rtol=1e-7)

# replicas stay bit-identical
W = torch.as_tensor(model.asignatures.X).cuda()
ref = W.clone()
dist.broadcast(ref, src=0)
assert torch.equal(W, ref)

# fp32 / tensor-core path at a size where it is taken, local shards given directly (shard_input=False)
import bench  # noqa: E402

D, k = 40_000, 8
lo, hi = bench.shard_bounds(D, world, rank)
Xl = bench.synth_rows(lo, hi, k).astype(np.float64)
W0, H0 = bench.init_rows(Xl, lo, k)
m3 = sal.models.KLNMF(n_signatures=k, init_method="custom", min_iterations=60, max_iterations=60, dtype="float32", math="tf32",
                      device=f"cuda:{local}", shard_input=False)
m3.fit(AnnData(Xl), init_kwargs={"signatures_mat": W0, "exposures_mat": H0})
kl_multi = m3.history["objective_function"][-1]
if rank == 0:
    Xf = bench.synth_rows(0, D, k).astype(np.float64)
    W0f, H0f = bench.init_rows(Xf, 0, k)
    from oracle import klnmf as oracle_klnmf

    Wo, Ho = W0f.T.copy(), H0f.T.copy()
    for _ in range(60):
        Wo, Ho = oracle_klnmf.update_WH(Xf.T, Wo, Ho)
    kl_ref = oracle_klnmf.kl_divergence(Xf.T, Wo, Ho)
    assert abs(kl_multi - kl_ref) / kl_ref < 1e-4, (kl_multi, kl_ref)
    A, B = m3.asignatures.X, Wo.T
    cos = np.sum(A * B, axis=1) / (np.linalg.norm(A, axis=1) * np.linalg.norm(B, axis=1))
    assert cos.min() >= 0.9999, cos
    print(f"multi-GPU ok: world {world}, fp64 trajectories match, CorrNMFDet matches the oracle, tf32 60-iteration KL {kl_multi:.3f} vs oracle {kl_ref:.3f}")
# the tf32 fit above ran on the persistent period kernel with the numerator exchange over NVLink INSIDE the kernel
# (sal_klnmf_period, n_ranks = world): replicas must hold bit-identical signatures
assert m3.launch_stats["driver"] == "persistent period kernel", m3.launch_stats
Wp = torch.as_tensor(np.asarray(m3.asignatures.X), device=f"cuda:{local}")
gathered = [torch.empty_like(Wp) for _ in range(world)]
dist.all_gather(gathered, Wp)
assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas must hold bit-identical signatures"
# the same fit on ONE rank holding everything (replica=True ignores the process group): equal to fp32 rounding
if rank == 0:
    solo = sal.models.KLNMF(n_signatures=k, init_method="custom", min_iterations=60, max_iterations=60, dtype="float32", math="tf32",
                            device=f"cuda:{local}", replica=True)
    solo.fit(AnnData(Xf), init_kwargs={"signatures_mat": W0f, "exposures_mat": H0f})
    assert np.allclose(np.asarray(solo.asignatures.X), np.asarray(m3.asignatures.X), rtol=1e-3, atol=1e-9)
    assert abs(solo.history["objective_function"][-1] - kl_multi) / kl_multi < 1e-6
# a convergence-driven fit (speculative periods past min_iterations) stops at the same iteration on every rank
W04, H04 = bench.init_rows(Xl, lo, 4)
mc2 = sal.models.KLNMF(n_signatures=4, init_method="custom", min_iterations=50, max_iterations=400, tol=1e-4, dtype="float32", math="tf32",
                       device=f"cuda:{local}", shard_input=False)
mc2.fit(AnnData(Xl), init_kwargs={"signatures_mat": W04, "exposures_mat": H04})
n_all = [None] * world
dist.all_gather_object(n_all, (mc2.n_iterations, len(mc2.history["objective_function"]), mc2.history["objective_function"][-1]))
assert len(set(n_all)) == 1 and 50 <= mc2.n_iterations < 400, n_all
dist.barrier()
dist.destroy_process_group()
