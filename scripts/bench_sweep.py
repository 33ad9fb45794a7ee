"""Model-selection sweep (BASELINE config 3: KLNMF k = 2..30 x 50 random restarts on 96 x 100k synthetic counts) on a bounded
sample: a few k values x 2 restarts x a fixed number of iterations through `salamander_b200.sweep.sweep_klnmf`, next to the
oracle (the reference's arithmetic, all host threads) timed for a few iterations at the same sizes.  Prints one JSON line.
oracle/ is used only as the CPU arm being timed."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from salamander_b200 import AnnData  # noqa: E402
from salamander_b200.sweep import error_curve, sweep_klnmf  # noqa: E402

D, n_it = int(os.environ.get("D", 100_000)), int(os.environ.get("ITERS", 500))
ks = [int(x) for x in os.environ.get("KS", "2,5,13,16,30").split(",")]
n_rs = int(os.environ.get("RESTARTS", 2))
# under torchrun the (k, seed) jobs are dealt round-robin to the ranks: replicas of X, no data-path collective
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = bench.synth_rows(0, D, 12)
adata = AnnData(X)
extra = {"init_device": True} if os.environ.get("INIT_DEVICE", "0") == "1" else {}
extra["device"] = f"cuda:{local}"
sweep_klnmf(adata, ks, n_restarts=max(1, world // len(ks) + 1), min_iterations=20, max_iterations=20, dtype="float32", math="tf32", **extra)  # warm-up: every k on every rank
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
table, best = sweep_klnmf(adata, ks, n_restarts=n_rs, min_iterations=n_it, max_iterations=n_it, dtype="float32", math="tf32", **extra)
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0
n_fits = len(table)
if world > 1:
    t = torch.tensor([gpu_s], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_s = float(t.item())
    if rank != 0:
        dist.barrier()
        os._exit(0)

from oracle.klnmf_mt import HostKLNMF  # noqa: E402

cpu_s_per_it = {}
if os.environ.get("NO_CPU", "0") != "1":  # (skipped on a many-GPU box: every second of host work there is charged for all GPUs)
    host = HostKLNMF(X.astype(np.float64))
    for k in (ks[0], ks[-1]):
        W0, H0 = bench.init_rows(X, 0, k)
        W, H = host.update_WH(W0, H0.copy())
        t0 = time.perf_counter()
        for _ in range(5):
            W, H = host.update_WH(W, H)
        cpu_s_per_it[k] = (time.perf_counter() - t0) / 5
    host.close()
cpu_mean = float(np.mean(list(cpu_s_per_it.values()))) if cpu_s_per_it else float('nan')
print(json.dumps({
    "workload": f"KLNMF sweep on synthetic 96 x {D}: k in {ks} x {n_rs} random restarts x {n_it} iterations, fp32 tensor-core path, {world} GPU(s), jobs round-robin over the ranks",
    "n_gpus": world,
    "init_device": bool(extra.get("init_device")), "fits": n_fits, "gpu_seconds": gpu_s, "gpu_fits_per_s": n_fits / gpu_s, "gpu_iterations_per_s": n_fits * n_it / gpu_s,
    "error_curve": {int(k): float(v) for k, v in error_curve(table).items()},
    "cpu_oracle_seconds_per_iteration": {int(k): v for k, v in cpu_s_per_it.items()}, "cpu_threads": os.cpu_count(),
    "full_sweep_extrapolation": {
        "fits": 29 * 50,
        "gpu_hours_one_gpu_at_this_rate_per_1000_iterations": 29 * 50 * 1000 / (n_fits * n_it / gpu_s) / 3600,
        "cpu_hours_per_1000_iterations": 29 * 50 * 1000 * cpu_mean / 3600,
    },
}))
if world > 1:
    dist.barrier()
    os._exit(0)
