"""
CorrNMFDet at scale (BASELINE config 5: 96 x 200k synthetic): milliseconds per iteration on the GPU, next to the
oracle (numpy + scipy Newton-CG, the reference's arithmetic) timed on a 2,000-sample slice and scaled linearly in D
(the per-sample Newton-CG loop dominates).  Prints one JSON line.  oracle/ is used only as the CPU arm being timed.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402

import os  # noqa: E402

import torch.distributed as dist  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:  # torchrun: samples sharded over the ranks (every rank holds the full host matrix and uploads its rows)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

D, k, m = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 5, 4
X = bench.synth_rows(0, D, k).astype(np.float64)
model = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, init_method="random", dtype="float64", device=f"cuda:{local}")
adata = AnnData(X)
model._setup_adata(adata)
np.random.seed(0)
model._initialize(None, {"seed": 0})
state0 = dict(W=np.array(model.asignatures.X), a=np.array(model.asignatures.obs["scalings"].values, dtype=float),
              b=np.array(adata.obs["scalings"].values, dtype=float), L=np.array(model.asignatures.obsm["embeddings"]),
              U=np.array(adata.obsm["embeddings"]), var=float(model.variance))
with model._resident():
    model._in_fit = True
    for _ in range(3):
        model._update_parameters(None)
    model.objective_function()  # (the constant sum lnGamma(1 + x) of the ELBO is computed once, here)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_it = 10
    e0.record()
    for _ in range(n_it):
        model._update_parameters(None)
    elbo = model.objective_function()
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / n_it
    model._in_fit = False

if rank != 0:
    dist.barrier()
    os._exit(0)
if os.environ.get("NO_CPU", "0") == "1":  # (many-GPU boxes: host seconds are charged for every GPU)
    print(json.dumps({"workload": f"CorrNMFDet k={k} dim={m} on synthetic 96 x {D}", "n_gpus": world, "gpu_ms_per_iteration": gpu_ms, "elbo": elbo}))
    if world > 1:
        dist.barrier()
    os._exit(0)
from oracle import corrnmf as oracle  # noqa: E402

Ds = 2000
Xs = X[:Ds]
t0 = time.perf_counter()
oracle.update_parameters(Xs, state0["W"], state0["a"], state0["b"][:Ds], state0["L"], state0["U"][:Ds], state0["var"])
cpu_s = time.perf_counter() - t0
print(json.dumps({"workload": f"CorrNMFDet k={k} dim={m} on synthetic 96 x {D}", "n_gpus": world, "gpu_ms_per_iteration": gpu_ms, "elbo": elbo,
                  "cpu_oracle_s_per_iteration_scaled": cpu_s * D / Ds, "cpu_sample": f"1 iteration on {Ds} samples: {cpu_s:.2f} s",
                  "speedup": cpu_s * D / Ds / (gpu_ms * 1e-3)}))
if world > 1:
    dist.barrier()
    os._exit(0)
