"""MvNMF at scale (BASELINE config 1's model on config 2's data shape): milliseconds per iteration on one B200 with the exact
FMA passes vs the tensor-core passes (math='tf32').  Prints one JSON line per mode."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402

D, k, n_it = int(os.environ.get("D", 1_000_000)), int(os.environ.get("K", 10)), int(os.environ.get("ITERS", 60))
X = bench.synth_rows(0, D, k)
W0, H0 = bench.init_rows(X, 0, k)
for math in ("fma", "tf32"):
    m = sal.models.MvNMF(n_signatures=k, init_method="custom", lam=1.0, delta=1.0, min_iterations=n_it, max_iterations=n_it,
                         dtype="float32", math=math)
    ad = AnnData(X)
    m._setup_adata(ad)
    m._initialize(None, {"signatures_mat": W0, "exposures_mat": H0})
    m._setup_fitting_parameters(None)
    with m._resident() as st:
        m._in_fit = True
        for _ in range(5):
            m._update_parameters(None)
        torch.cuda.synchronize()
        n0 = st.ws.launches
        t0 = time.perf_counter()
        for _ in range(n_it):
            m._update_parameters(None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        obj = m.objective_function()
        launches = (st.ws.launches - n0) / n_it
        m._in_fit = False
    print(json.dumps({"workload": f"MvNMF k={k} lam=1 delta=1 on synthetic 96 x {D}, fp32", "math": math, "ms_per_iteration": dt / n_it * 1e3,
                      "launches_per_iteration": launches, "objective": obj}))
