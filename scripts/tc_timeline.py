"""Diagnostics (GPU): print the per-role clock64 timeline CTA 0 of the tensor-core pass records (sal_set_debug_buffer)."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from salamander_b200._device import PASS_UPDATE_H, PASS_WNUM, Workspace

dev = torch.device("cuda:0")
D, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 20
gen = torch.Generator(device=dev).manual_seed(0)
W = torch.rand((k, 96), generator=gen, device=dev) + 0.01
W /= W.sum(1, keepdim=True)
H = torch.rand((D, k), generator=gen, device=dev) * 400 + 1
X = torch.poisson(H @ W, generator=gen).clamp_min(1e-7)
ws = Workspace(96, D, k, torch.float32, dev, math="tf32")
Wnum = torch.empty_like(W)
for _ in range(3):
    ws.klnmf_pass(X, W, H, PASS_UPDATE_H | PASS_WNUM, H_out=H, Wnum=Wnum)
dbg = torch.zeros(32768, dtype=torch.float32, device=dev)
ws.set_debug_buffer(dbg)
ws.klnmf_pass(X, W, H, PASS_UPDATE_H | PASS_WNUM, H_out=H, Wnum=Wnum)
torch.cuda.synchronize()
ws.set_debug_buffer(None)
t = dbg.view(torch.int32).cpu().numpy().astype(np.int64)[128 * 96 + 128 * 32 :][: 4 * 48 * 8].reshape(4, 48, 8) & 0xFFFFFFFF
t0 = t[3, 0, 0]
names = ["start", "rready>", "P0nxt>", "whfull", "E1done", "shtfree", "sts+wst", "hnfull"]
print("WG timeline (clk since first TMA issue): tile | " + " ".join(n.rjust(8) for n in names) + " | next-start")
for i in range(16):
    g = i & 1
    row = (t[g, i] - t0).tolist()
    print(f"WG{g} tile {i:2d} | " + " ".join(f"{v:8d}" for v in row))
print("MMA: tile | hready-seen  G1-issued  rready-seen  G2G3-issued")
for i in range(16):
    print(f"   {i:2d} | {t[2, i, 0] - t0:8d} {t[2, i, 2] - t0:8d} {t[2, i, 1] - t0:8d} {t[2, i, 3] - t0:8d}")
print("kernel entry, setup done, kernel exit (clk relative to the first TMA issue):", int(t[3, 0, 1] - t0), int(t[3, 0, 2] - t0), int(t[3, 0, 3] - t0))
print("TMA: tile | issue")
print("  ", [int(t[3, i, 0] - t0) for i in range(16, 32)])
print("  ", [int(t[3, i, 0] - t0) for i in range(16)])
