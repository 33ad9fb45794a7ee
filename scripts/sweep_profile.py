"""Host-side profile of a bounded k-sweep (diagnostic): where do the milliseconds per fit go?"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402
from salamander_b200.sweep import sweep_klnmf  # noqa: E402

D, n_it = int(os.environ.get("D", 100_000)), int(os.environ.get("ITERS", 200))
ks = [int(x) for x in os.environ.get("KS", "2,5,13,30").split(",")]
X = bench.synth_rows(0, D, 12)
adata = AnnData(X)
kw = dict(min_iterations=n_it, max_iterations=n_it, dtype="float32", math="tf32", init_device=True)
sweep_klnmf(adata, [4], n_restarts=1, **{**kw, "min_iterations": 20, "max_iterations": 20})
sweep_klnmf(adata, ks, n_restarts=2, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
table, _ = sweep_klnmf(adata, ks, n_restarts=2, **kw)
torch.cuda.synchronize()
t = time.perf_counter() - t0
print(f"{len(table)} fits in {t * 1e3:.1f} ms = {t / len(table) * 1e3:.2f} ms per fit, {len(table) * n_it / t:.0f} iterations / s")
pr = cProfile.Profile()
pr.enable()
sweep_klnmf(adata, ks, n_restarts=2, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
