#!/usr/bin/env python
"""
bench.py -- KLNMF iterations/s at 96 x 1M, k=20 on 1/2/4/8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1, one rank per GPU

A "step" is one KLNMF iteration (the joint multiplicative update of W and H, reference
_utils_klnmf.py:281-361) over the whole synthetic count matrix; every ``conv_test_freq``-th
iteration also evaluates the KL objective and brings it to the host, exactly as
``SignatureNMF.fit`` does (reference signature_nmf.py:365-380).  The samples are sharded over
the ranks (strong scaling: the job is always 96 x 1M), H stays local and the 96 x k numerator is
all-reduced each iteration.

Printed JSON (rank 0, one line): see the task contract; ``value`` is device-resident throughput,
``e2e`` goes through ``KLNMF.fit(adata)`` with HOST arrays (pinned), upload and download inside
the timed region, ``roofline`` is the fused pass kernel against MEASURED_PEAKS.json, and
``cpu_baseline`` is the multi-threaded numpy port of the reference (oracle/klnmf_mt.py) on this
box's cores.  oracle/ is used here ONLY as the CPU arm that is being timed.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20240518
V = 96
CHUNK = 62_500  # generation granularity: results do not depend on the number of ranks
EPS = float(np.finfo(np.float32).eps)


# --------------------------------------------------------------------------------------------
# synthetic Poisson SBS-96 counts
# --------------------------------------------------------------------------------------------
def true_signatures(k_true: int) -> np.ndarray:
    rng = np.random.default_rng(SEED)
    return rng.dirichlet(0.5 * np.ones(V), size=k_true)  # [k][V]


def synth_rows(lo: int, hi: int, k_true: int, out: np.ndarray | None = None) -> np.ndarray:
    """Rows [lo, hi) of the D x 96 count matrix as float32, clipped to EPSILON (signature_nmf.py:281)."""
    Wt = true_signatures(k_true)
    X = np.empty((hi - lo, V), dtype=np.float32) if out is None else out
    c0, c1 = lo // CHUNK, (hi - 1) // CHUNK if hi > lo else -1
    for c in range(c0, c1 + 1):
        rng = np.random.default_rng(SEED + 1000 + c)
        burden = np.exp(rng.normal(np.log(5000.0), 0.8, size=CHUNK))
        act = rng.dirichlet(0.3 * np.ones(k_true), size=CHUNK)
        lam = (burden[:, None] * act) @ Wt
        cnt = rng.poisson(lam).astype(np.float32)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        X[a - lo : b - lo] = cnt[a - c * CHUNK : b - c * CHUNK]
    np.maximum(X, np.float32(EPS), out=X)
    return X


def init_rows(X: np.ndarray, lo: int, k: int):
    """init_random semantics (reference initialization/methods.py:89-109), generated per chunk so that the
    start is independent of the rank count; then normalise / clip as initialize_mat does (:116-118)."""
    rng = np.random.default_rng(SEED + 1)
    W0 = rng.dirichlet(np.ones(V), size=k)
    W0 = np.maximum(W0 / W0.sum(axis=1, keepdims=True), EPS)
    H0 = np.empty((X.shape[0], k), dtype=np.float64)
    hi = lo + X.shape[0]
    for c in range(lo // CHUNK, ((hi - 1) // CHUNK if hi > lo else -1) + 1):
        rng = np.random.default_rng(SEED + 5000 + c)
        d = rng.dirichlet(np.ones(k), size=CHUNK)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        H0[a - lo : b - lo] = d[a - c * CHUNK : b - c * CHUNK]
    H0 *= X.sum(axis=1, dtype=np.float64)[:, None]
    np.maximum(H0, EPS, out=H0)
    return W0, H0


def shard_bounds(n: int, world: int, rank: int):
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# --------------------------------------------------------------------------------------------
# clocks during the timed region (NVML poller thread)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    NOTED = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in {**self.BAD, **self.NOTED}.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "n_samples": len(self.samples),
        }


# --------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), bounded sample
# --------------------------------------------------------------------------------------------
def cpu_iterations_per_s(D_total: int, k: int, n_steps: int, warmup: int, budget_s: float, conv_test_freq: int):
    """Times the multi-threaded numpy port on a sample of the SAME synthetic matrix (its first D_s rows) and
    scales linearly in D to the full 96 x D_total job.  Returns (it/s at full size, description, threads)."""
    from oracle.klnmf_mt import HostKLNMF, n_host_threads

    threads = n_host_threads()
    D_s = min(D_total, 100_000)
    X = synth_rows(0, D_s, k).astype(np.float64)
    W, H = init_rows(X, 0, k)
    host = HostKLNMF(X, threads)
    t0 = time.perf_counter()
    W, H = host.update_WH(W, H)
    t_probe = time.perf_counter() - t0
    # shrink the sample if n_steps iterations of it would not fit the time budget
    per_step_budget = budget_s / max(1, n_steps + warmup)
    if t_probe > per_step_budget and D_s > 5_000:
        D_s = max(5_000, int(D_s * per_step_budget / t_probe) // 1000 * 1000)
        host.close()
        X = np.ascontiguousarray(X[:D_s])
        W, H = init_rows(X, 0, k)
        host = HostKLNMF(X, threads)
    for _ in range(warmup):
        W, H = host.update_WH(W, H)
    t0 = time.perf_counter()
    for it in range(1, n_steps + 1):
        W, H = host.update_WH(W, H)
        if it % conv_test_freq == 0:
            host.kl_divergence(W, H)
    dt = time.perf_counter() - t0
    host.close()
    its = n_steps / dt * (D_s / D_total)
    sample = (
        f"{n_steps} update_WH iterations (+ kl_divergence every {conv_test_freq}) on the first {D_s} of {D_total} "
        f"synthetic samples, float64 numpy port with {threads} threads; it/s scaled by {D_s}/{D_total} (cost is linear in D)"
    )
    return its, sample, threads, dt / n_steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    its, sample, threads, ms = cpu_iterations_per_s(
        args.samples, args.k, args.steps, args.warmup, budget_s=150.0, conv_test_freq=args.conv_test_freq
    )
    line = {
        "impl": "reference",
        "metric": "KLNMF iterations/s at 96x1M k=20",
        "value": its,
        "unit": "iterations/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 / its,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": its, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": its, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": f"KLNMF k={args.k} on synthetic Poisson SBS-96 counts, 96 x {args.samples} samples (BASELINE configs[2])",
        "n_features": V,
        "n_samples": args.samples,
        "n_signatures": args.k,
        "conv_test_freq": args.conv_test_freq,
        "parallelism": f"samples sharded over {world} GPU(s); 96 x k numerator all-reduced per iteration",
    }


# --------------------------------------------------------------------------------------------
# this repo's CUDA path
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import salamander_b200 as sal
    from salamander_b200 import AnnData

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    D, k = args.samples, args.k
    lo, hi = shard_bounds(D, world, rank)
    t0 = time.perf_counter()
    # pinned host buffers: the e2e leg uploads from them
    X_pin = torch.empty((hi - lo, V), dtype=torch.float32).pin_memory()
    X_host = X_pin.numpy()
    synth_rows(lo, hi, k, out=X_host)
    W0, H0 = init_rows(X_host, lo, k)
    t_gen = time.perf_counter() - t0

    def make_model(n_iter):
        return sal.models.KLNMF(
            n_signatures=k,
            init_method="custom",
            min_iterations=n_iter,
            max_iterations=n_iter,
            conv_test_freq=args.conv_test_freq,
            dtype="float32",
            math=args.math,
            device=dev,
            shard_input=False,
        )

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") ---------------------------------------------
    model = make_model(args.steps)
    adata = AnnData(X_host)
    model._setup_adata(adata)
    model._initialize(None, {"signatures_mat": W0, "exposures_mat": H0})
    model._setup_fitting_parameters(None)
    model._to_device()
    st = model._dev

    def run_loop(n_iter):
        """``n_iter`` iterations of the model's own device-side fit driver (KLNMF._fit_loop): updates, the objective
        every conv_test_freq iterations and its read-back to the host, exactly what fit() runs after the upload."""
        model.min_iterations = model.max_iterations = n_iter
        return model._fit_loop(None, 0, 10**9)

    # warm-up: at least W iterations; enough periods for the fit driver to have captured its CUDA graphs
    n_warm = max(args.warmup, 6 * args.conv_test_freq)
    run_loop(n_warm)
    barrier()
    launches0 = st.ws.launches
    with ClockSampler(local_rank) as clocks:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        of_values, n_done = run_loop(args.steps)
        ev1.record()
        barrier()
    assert n_done == args.steps
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    its = args.steps / (elapsed_ms * 1e-3)
    launch_stats = dict(model.launch_stats)
    launches_graph = st.ws.launches - launches0  # launches issued while capturing / running eagerly

    # roofline leg: the same loop run eagerly with CUDA events around every UPDATE_H | WNUM pass kernel (events cannot
    # be recorded inside a replayed graph), on the stream the kernel is launched on
    n_roof = min(args.steps, 100)
    st.ws.set_timing(True)
    launches1 = st.ws.launches
    run_loop(n_roof)
    kernel_ms_total, n_timed = st.ws.pass_timing()
    st.ws.set_timing(False)
    launches_per_step = (st.ws.launches - launches1) / n_roof
    launches = int(round(launches_per_step * args.steps))
    kernel_ms_bracketed = kernel_ms_total / n_timed if n_timed else float("nan")
    final_kl = model.objective_function()
    model._to_host()

    # The per-launch event pairs above put an event record (wait-for-idle + timestamp write) on both sides of every
    # kernel, which adds several microseconds to a ~90 us kernel.  The figure the roofline uses is therefore the average
    # over a chain of launches of the SAME kernel on the same operands between ONE pair of events: the streaming kernel
    # alone (SAL_PASS_PARTIALS_ONLY: the 6 us reduction kernel is not launched in between), H ping-ponging between two
    # buffers so that every launch reads X and H from HBM and writes H (544 MB per launch > L2).
    from salamander_b200 import _lib as sal_lib

    chain_flags = sal_lib.PASS_UPDATE_H | sal_lib.PASS_WNUM | sal_lib.PASS_PARTIALS_ONLY
    H_a, H_b = st.H.clone(), torch.empty_like(st.H)
    n_chain = 50

    def chain(n):
        nonlocal H_a, H_b
        for _ in range(n):
            st.ws.klnmf_pass(st.X, st.W, H_a, chain_flags, H_out=H_b)
            H_a, H_b = H_b, H_a

    chain(4)
    torch.cuda.synchronize()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    chain(n_chain)
    ev3.record()
    torch.cuda.synchronize()
    kernel_ms = ev2.elapsed_time(ev3) / n_chain
    del H_a, H_b
    model._release_device()

    # ---- end to end through the public API: KLNMF.fit(adata) with host arrays --------------
    e2e_model = make_model(args.steps)
    H0_pin = torch.from_numpy(H0).pin_memory().numpy()
    fit_times = []
    for rep in range(6):  # repetition 0 warms the allocator caches; the median of the other five is reported
        adata2 = AnnData(X_host)
        barrier()
        t0 = time.perf_counter()
        e2e_model.fit(adata2, init_kwargs={"signatures_mat": W0, "exposures_mat": H0_pin})
        torch.cuda.synchronize()
        if rep:
            fit_times.append(time.perf_counter() - t0)
    t_fit = float(np.median(fit_times))
    t_fit = max_over_ranks(t_fit)
    h2d = e2e_model.transfer_bytes["h2d"]
    d2h = e2e_model.transfer_bytes["d2h"]
    if world > 1:
        tb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(tb)
        h2d, d2h = tb.tolist()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        D_local = hi - lo
        alg_bytes = V * D_local * 4 + 2 * k * D_local * 4  # read X once, read H once, write H once (fp32)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("n_samples_per_gpu") == D_local and tj.get("k") == k and tj.get("math") == args.math:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "KLNMF iterations/s at 96x1M k=20",
            "value": its,
            "unit": "iterations/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32" if args.math == "fma" else "f32 (tf32 tensor-core contractions, f32 accumulate)",
            "data": "synthetic",
            "config": {
                **workload_config(args, world),
                "math": args.math,
                "l2": (
                    f"inputs {alg_bytes / 1e6:.0f} MB per GPU per iteration "
                    + ("> 126 MB L2, no flush needed" if alg_bytes > 1.5 * 126e6 else "fit in L2 (strong-scaling shard); not flushed")
                ),
                "timing": "CUDA events on the launch stream around KLNMF._fit_loop (updates + objective read-back every conv_test_freq iterations; "
                + ("CUDA-graph replays" if launch_stats.get("graphs") else "eager launches, one speculative period ahead")
                + "), max over ranks",
                "warmup_iterations_run": n_warm,
                "fit_driver": launch_stats,
                "final_kl": final_kl,
                "data_generation_s": t_gen,
            },
            "clocks": clocks.summary(),
            "e2e": {
                "value": args.steps / t_fit,
                "unit": "iterations/s",
                "h2d_bytes_per_step": h2d / args.steps,
                "d2h_bytes_per_step": d2h / args.steps,
                "what": f"KLNMF.fit(adata) of {args.steps} iterations from pinned host arrays: upload of X, W0, H0 and download of W, H inside the timed region (wall clock, max over ranks)",
                "seconds": t_fit,
                "seconds_all": fit_times,
            },
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm",
                "kernel": "klnmf_pass (UPDATE_H|WNUM)",
                "achieved": achieved,
                "peak": peak_gbs,
                "unit": "GB/s",
                "frac": achieved / peak_gbs,
                "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": kernel_ms,
                "n_launches_timed": n_chain,
                "how": f"one CUDA event pair around {n_chain} back-to-back launches of the pass kernel alone (same X, W; H ping-pong; reduction kernel skipped) on the launch stream, right after the timed run",
                "kernel_ms_event_pair_per_launch": kernel_ms_bracketed,
                "how_event_pair_per_launch": f"CUDA events around each pass kernel during {n_roof} eager iterations of the fit loop ({n_timed} launches); includes the two event records' wait-for-idle",
                "frac_floor_from_whole_step": alg_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9 / peak_gbs,
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            cits, sample, threads, _ = cpu_iterations_per_s(D, k, 20, 1, budget_s=25.0, conv_test_freq=args.conv_test_freq)
            line["cpu_baseline"] = {"value": cits, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL collectives keep the communicator busy at teardown: drop them, make sure every
        # rank is done, and leave without running the (occasionally hanging) communicator destructors
        del model, e2e_model
        import gc

        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--samples", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--conv-test-freq", type=int, default=10)
    ap.add_argument("--math", choices=["fma", "tf32"], default="tf32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
