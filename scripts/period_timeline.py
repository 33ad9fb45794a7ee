"""Timeline of the persistent period kernel (diagnostic): %globaltimer stamps of one epilogue thread of CTA 1 per update.

slots: 0 sweep start | 1 tile loop done | 2 all MMAs retired | 3 partial published | 4 slice reduced, totals published |
       5 all totals polled | 6 W operands rewritten.   D, K, L from the environment."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salamander_b200._device import Workspace  # noqa: E402

D, k, L = int(os.environ.get("D", 1_000_000)), int(os.environ.get("K", 20)), int(os.environ.get("L", 10))
every = int(os.environ.get("EVERY", 0))
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
W = torch.rand((k, 96), generator=gen, device=dev) + 0.01
W /= W.sum(1, keepdim=True)
H = torch.rand((D, k), generator=gen, device=dev) * 400 + 1
X = torch.poisson(H @ W, generator=gen).clamp_min(1e-7)
ws = Workspace(96, D, k, torch.float32, dev, math="tf32_always")
W2, H2 = torch.empty_like(W), torch.empty_like(H)
n_obj = (-(-L // every) if every else 0)
objs = torch.zeros(max(n_obj, 1), dtype=torch.float64, device=dev)
G = min(148, -(-D // 128))
tl = torch.zeros(8 * (L + 1) + 4 * L * G, dtype=torch.int64, device=dev)
for rep in range(3):
    ws.klnmf_period(X, W, W2, H, H2, 0, True, L, every, False, objectives=objs if n_obj else None)
ws.set_debug_buffer(tl)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
ws.klnmf_period(X, W, W2, H, H2, 0, True, L, every, False, objectives=objs if n_obj else None)
ev1.record()
torch.cuda.synchronize()
ws.set_debug_buffer(None)
raw = tl.cpu().numpy()
t = raw[: 8 * (L + 1)].reshape(L + 1, 8)[:L, :8].astype(np.float64)
ct = raw[8 * (L + 1):].reshape(L, G, 4).astype(np.float64)
t0 = t[0, 0]
print(f"D={D} k={k} L={L} every={every}: kernel {ev0.elapsed_time(ev1) * 1e3:.1f} us = {ev0.elapsed_time(ev1) * 1e3 / L:.2f} us / update")
print("update  start  | tiles   mma-done  partial  stageA   stageB   W-done  loopend | total (us)")
for u in range(L):
    r = t[u]
    d = np.diff(r) / 1e3
    nxt = (t[u + 1, 0] - r[0]) / 1e3 if u + 1 < L else float("nan")
    print(f"{u:4d} {(r[0] - t0) / 1e3:9.2f} | " + " ".join(f"{x:8.2f}" for x in d) + f" | {nxt:8.2f}")
print("per-CTA skew (us, relative to the first CTA to finish its tiles): tiles-done max | all-MMA-done max | slice published min..max | W epilogue done min..max")
for u in range(L):
    b = ct[u, :, 0].min()
    print(f"{u:4d}  {(ct[u, :, 0].max() - b) / 1e3:7.2f} | {(ct[u, :, 1].max() - b) / 1e3:7.2f} | {(ct[u, :, 2].min() - b) / 1e3:7.2f} .. {(ct[u, :, 2].max() - b) / 1e3:7.2f} | "
          f"{(ct[u, :, 3].min() - b) / 1e3:7.2f} .. {(ct[u, :, 3].max() - b) / 1e3:7.2f}   slowest CTA (MMA done): {int(ct[u, :, 1].argmax())}")
ws.close()
