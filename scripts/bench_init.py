"""NNDSVD initialisation at the bench workload (96 x 1M, k = 20): scikit-learn's randomized SVD on the host (the reference's
route) vs the Gram-matrix SVD on the device.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from salamander_b200.initialization.device_nndsvd import init_nndsvd_device  # noqa: E402
from salamander_b200.initialization.methods import init_nndsvd  # noqa: E402

D, k = int(os.environ.get("D", 1_000_000)), 20
X = bench.synth_rows(0, D, k).astype(np.float64)
dev = torch.device("cuda:0")
init_nndsvd_device(X[:10_000], k, seed=0, device=dev)  # context / cuBLAS warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
s_dev, e_dev = init_nndsvd_device(X, k, seed=0, device=dev)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
s_ref, e_ref = init_nndsvd(X, k, seed=0)
t_host = time.perf_counter() - t0
print(json.dumps({
    "workload": f"init_method='nndsvd', 96 x {D}, k={k}, float64 host matrix in, host factors out",
    "device_seconds": t_dev, "host_sklearn_seconds": t_host,
    "max_abs_diff_signatures_over_scale": float(np.abs(s_dev - s_ref).max() / np.abs(s_ref).max()),
    "max_abs_diff_exposures_over_scale": float(np.abs(e_dev - e_ref).max() / np.abs(e_ref).max()),
    "zero_pattern_differences": int(((s_dev == 0) != (s_ref == 0)).sum() + ((e_dev == 0) != (e_ref == 0)).sum()),
}))
