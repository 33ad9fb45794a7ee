"""Probe for ncu: a few launches of the period kernel whose every update carries the KL objective (EVERY=1) or none (EVERY=0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salamander_b200._device import Workspace  # noqa: E402

D, k, L = int(os.environ.get("D", 1_000_000)), int(os.environ.get("K", 20)), int(os.environ.get("L", 4))
every = int(os.environ.get("EVERY", 1))
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
W = torch.rand((k, 96), generator=gen, device=dev) + 0.01
W /= W.sum(1, keepdim=True)
H = torch.rand((D, k), generator=gen, device=dev) * 400 + 1
X = torch.poisson(H @ W, generator=gen).clamp_min(1e-7)
ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
W2, H2 = torch.empty_like(W), torch.empty_like(H)
n_obj = -(-L // every) if every else 0
objs = torch.zeros(max(n_obj, 1), dtype=torch.float64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    ev0.record()
    ws.klnmf_period(X, W, W2, H, H2, 0, True, L, every, False, objectives=objs if n_obj else None)
    ev1.record()
    torch.cuda.synchronize()
print(f"D={D} k={k} L={L} every={every}: {ev0.elapsed_time(ev1) * 1e3 / L:.2f} us / update")
ws.close()
