"""
BASELINE configs 0 and 1 (the reference's own CPU-runnable cases) on the GPU next to the oracle on the host:
  C1  KLNMF n_signatures=5 on PCAWG breast SBS (96 x 192), default stopping rule, fp64
  C2  MvNMF n_signatures=10 on the same data, lam = delta = 1, 2,000 iterations, fp64
Prints one JSON line per config.  oracle/ is used only as the CPU arm being timed.
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, ".")
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402
from oracle import EPSILON, klnmf as oklnmf, mvnmf as omvnmf  # noqa: E402

counts = pd.read_csv(os.path.join("salamander_b200", "data", "pcawg_breast_sbs.csv"), index_col=0).T


def timed_fit(model, **kw):
    model.fit(AnnData(counts), **kw)  # warm-up (kernel variants, graphs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.fit(AnnData(counts), **kw)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


m = sal.models.KLNMF(n_signatures=5, init_method="random", dtype="float64")
t = timed_fit(m, init_kwargs={"seed": 0})
X = counts.values.astype(float).clip(EPSILON)
np.random.seed(0)
from salamander_b200.initialization.initialize import initialize_mat  # noqa: E402

W0, H0 = initialize_mat(X, 5, "random", seed=0)
t0 = time.perf_counter()
_, _, n_cpu, hist = oklnmf.fit_klnmf(X.T, W0.T, H0.T)
t_cpu = time.perf_counter() - t0
print(json.dumps({"config": "C1 KLNMF k=5 PCAWG 96x192 fp64, default stopping", "gpu_iterations": m.n_iterations, "gpu_seconds": t,
                  "gpu_it_per_s": m.n_iterations / t, "final_kl": m.history["objective_function"][-1],
                  "cpu_oracle_iterations": n_cpu, "cpu_oracle_seconds": t_cpu, "cpu_oracle_it_per_s": n_cpu / t_cpu, "cpu_final_kl": hist[-1]}))

m2 = sal.models.MvNMF(n_signatures=10, init_method="random", min_iterations=2000, max_iterations=2000, dtype="float64")
t2 = timed_fit(m2, init_kwargs={"seed": 0})
W0, H0 = initialize_mat(X, 10, "random", seed=0)
t0 = time.perf_counter()
res = omvnmf.fit_mvnmf(X.T, W0.T, H0.T, lam=1.0, delta=1.0, min_iterations=300, max_iterations=300)
t_cpu2 = (time.perf_counter() - t0) / 300 * 2000
print(json.dumps({"config": "C2 MvNMF k=10 PCAWG 96x192 fp64, 2000 iterations", "gpu_seconds": t2, "gpu_it_per_s": 2000 / t2,
                  "final_objective": m2.history["objective_function"][-1], "cpu_oracle_seconds_scaled": t_cpu2,
                  "cpu_oracle_it_per_s": 2000 / t_cpu2}))
