"""Fixed costs of a launch of the persistent period kernel (diagnostic): device time of one launch as a function of the number
of updates L, with and without the fused / trailing objective sweeps, at the 8-GPU shard size and at 1M samples.
T(L) = a + b L: `a` is what a launch costs besides its updates (cooperative launch, prologue, drain), `b` the time per update.

    D="125000 1000000" K=20 python scripts/period_fixed_costs.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salamander_b200._device import Workspace  # noqa: E402

k = int(os.environ.get("K", 20))
dev = torch.device("cuda", 0)


def timed(fn, reps=7):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(out))


for D in [int(x) for x in os.environ.get("D", "125000 1000000").split()]:
    gen = torch.Generator(device=dev).manual_seed(0)
    W = torch.rand((k, 96), generator=gen, device=dev) + 0.01
    W /= W.sum(1, keepdim=True)
    H = torch.rand((D, k), generator=gen, device=dev) * 400 + 1
    X = torch.poisson(H @ W, generator=gen).clamp_min(1e-7)
    ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
    W2, H2 = torch.empty_like(W), torch.empty_like(H)
    objs = torch.zeros(64, dtype=torch.float64, device=dev)
    print(f"D = {D}, k = {k}")
    rows = {}
    for name, every, fin in (("plain", 0, False), ("KL every 10", 10, False), ("KL every 10 + trailing objective", 10, True), ("trailing objective only", 0, True)):
        ts = []
        Ls = (1, 2, 5, 10, 20, 40)
        for L in Ls:
            fn = lambda: ws.klnmf_period(X, W, W2, H, H2, 0, True, L, every, fin, objectives=objs)  # noqa: E731
            fn()
            ts.append(timed(fn))
        b, a = np.polyfit(Ls, ts, 1)
        rows[name] = ts
        print(f"  {name:36s}: " + "  ".join(f"L={L}: {t:8.1f}" for L, t in zip(Ls, ts)) + f"   | fit: {a:6.1f} us + {b:6.2f} us / update")
    # an empty kernel launched the same way (cooperative, one CTA per SM) would cost: approximated by L = 0 + trailing objective
    fn = lambda: ws.klnmf_period(X, W, W2, H, H2, 0, True, 0, 0, True, objectives=objs)  # noqa: E731
    try:
        fn()
        print(f"  objective-only launch (L = 0): {timed(fn):8.1f} us")
    except Exception as exc:
        print("  objective-only launch not supported:", exc)
    ws.close()
    del X, H, H2
