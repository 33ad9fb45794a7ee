"""Probe for ncu: a few CorrNMFDet iterations at scale (synthetic 96 x D, k = 5, dim = 4, float64)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402

D, k, m, n_it = int(os.environ.get("D", 200_000)), 5, 4, int(os.environ.get("ITERS", 4))
X = bench.synth_rows(0, D, k).astype(np.float64)
model = sal.models.CorrNMFDet(n_signatures=k, dim_embeddings=m, init_method="random", dtype="float64")
ad = AnnData(X)
model._setup_adata(ad)
np.random.seed(0)
model._initialize(None, {"seed": 0})
with model._resident():
    model._in_fit = True
    for _ in range(2):
        model._update_parameters(None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_it + 1)]
    evs[0].record()
    for i in range(n_it):
        model._update_parameters(None)
        evs[i + 1].record()
    torch.cuda.synchronize()
    print(f"CorrNMFDet 96 x {D}: {(time.perf_counter() - t0) / n_it * 1e3:.3f} ms / iteration")
    print("per iteration (device, ms):", " ".join(f"{evs[i].elapsed_time(evs[i + 1]):.3f}" for i in range(n_it)))
    model._in_fit = False
