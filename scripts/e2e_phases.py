"""Where does KLNMF.fit(adata) spend its wall-clock time at the bench workload?  (diagnostic; prints one JSON line)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from salamander_b200._anndata import AnnData  # noqa: E402
from salamander_b200.models import KLNMF  # noqa: E402


def main():
    D = int(os.environ.get("D", 1_000_000))
    k, steps = 20, int(os.environ.get("STEPS", 500))
    # under torchrun: every rank fits its shard of the samples (as bench.py's e2e leg does)
    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = bench.shard_bounds(D, world, rank)
    X_host = torch.empty((hi - lo, 96), dtype=torch.float32).pin_memory().numpy()
    bench.synth_rows(lo, hi, k, out=X_host)
    W0, H0 = bench.init_rows(X_host, lo, k)
    H0_pin = torch.from_numpy(H0).pin_memory().numpy()
    out = []
    for rep in range(int(os.environ.get('REPS', 8))):
        m = KLNMF(n_signatures=k, init_method="custom", dtype="float32", math="tf32", min_iterations=steps,
                  max_iterations=steps, shard_input=False, device=f"cuda:{local}")
        m.use_graphs = {'1': True, '0': False}.get(os.environ.get('GRAPHS', 'auto'), 'auto')
        m.profile_phases = rep >= int(os.environ.get('PROFILE_FROM', 6))
        ad = AnnData(X_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.fit(ad, init_kwargs={"signatures_mat": W0, "exposures_mat": H0_pin})
        torch.cuda.synchronize()
        out.append({"rep": rep, "total_s": time.perf_counter() - t0, "phases": getattr(m, "phase_seconds", None)})
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
