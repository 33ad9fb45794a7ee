"""MvNMF at 96 x D, k = 10, float32 / tf32: the fit loop's device time per iteration with the run-ahead driver and with the plain
host loop (one host decision per line-search trial).  D, ITERS from the environment."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import salamander_b200 as sal  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402

D, k, n_it = int(os.environ.get("D", 1_000_000)), int(os.environ.get("K", 10)), int(os.environ.get("ITERS", 40))
X = bench.synth_rows(0, D, k)
W0, H0 = bench.init_rows(X, 0, k)
for run_ahead in (True, False):
    m = sal.models.MvNMF(n_signatures=k, init_method="custom", lam=1.0, delta=1.0, min_iterations=n_it, max_iterations=n_it,
                         dtype="float32", math="tf32")
    m.run_ahead = run_ahead
    m._setup_adata(AnnData(X))
    m._initialize(None, {"signatures_mat": W0, "exposures_mat": H0})
    m._setup_fitting_parameters(None)
    m._to_device()
    m._in_fit = True
    m._fit_loop(None, 0, 10**9)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = m._dev.ws.launches
    e0.record()
    of, n = m._fit_loop(None, 0, 10**9)
    e1.record()
    torch.cuda.synchronize()
    print(f"run_ahead={run_ahead}: {e0.elapsed_time(e1) / n:.4f} ms / iteration over {n} iterations (objective every {m.conv_test_freq}), "
          f"{(m._dev.ws.launches - n0) / n:.1f} launches / iteration, final objective {of[-1]:.6e}, stats {getattr(m, 'launch_stats', None)}")
    m._in_fit = False
    m._release_device()
