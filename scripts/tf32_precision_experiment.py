"""
Experiment (GPU, torch only -- not product code): which of the three contractions of the KL-NMF joint
update tolerate 1 x tf32 operands?  Emulates tf32 round-to-nearest on fp64 tensors and runs the
multiplicative updates from identical starts; prints the relative KL gap and the minimum signature
cosine against the exact fp64 run.  Used to choose the precision recipe of the tcgen05 pass
(DESIGN.md "tf32 recipe").
"""
import sys
import torch

EPS = 1.1920928955078125e-07
dev = torch.device("cuda:0")


def tf32(x):
    b = x.float().view(torch.int32)
    b = (b + 0x1000) & ~0x1FFF
    return b.view(torch.float32).double()


def step(X, W, H, mode):
    # W [k][V], H [D][k]
    g1 = {"exact": lambda: H @ W, "tf32": lambda: tf32(H) @ tf32(W)}
    WH = (H @ W) if mode["g1"] == "exact" else (tf32(H) @ tf32(W))
    R = X / WH
    if mode.get("fp32r"):
        R = R.float().double()
    Rt = tf32(R) if mode["r"] == "tf32" else R
    Wg2 = tf32(W) if mode["g2w"] == "tf32" else W
    Hg3 = tf32(H) if mode["g3h"] == "tf32" else H
    Rg3 = Rt if mode.get("g3r", "tf32") == "tf32" else R
    Hn = Rt @ Wg2.T
    Wn = Hg3.T @ Rg3
    Wnew = W * Wn
    Wnew = (Wnew / Wnew.sum(1, keepdim=True)).clamp_min(EPS)
    Hnew = (H * Hn).clamp_min(EPS)
    return Wnew, Hnew


def kl(X, W, H):
    WH = H @ W
    return float((X * torch.log(X / WH) - X + WH).sum())


def run(D, k, iters, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    Wt = torch.distributions.Dirichlet(torch.full((96,), 0.5, device=dev)).sample((k,)).double()
    burden = torch.exp(torch.randn(D, generator=g, device=dev, dtype=torch.float64) * 0.8 + 8.5)
    act = torch.distributions.Dirichlet(torch.full((k,), 0.3, device=dev)).sample((D,)).double()
    X = torch.poisson((burden[:, None] * act) @ Wt, generator=g).clamp_min(EPS)
    W0 = torch.distributions.Dirichlet(torch.ones(96, device=dev)).sample((k,)).double().clamp_min(EPS)
    H0 = (X.sum(1, keepdim=True) * torch.distributions.Dirichlet(torch.ones(k, device=dev)).sample((D,)).double()).clamp_min(EPS)
    modes = {
        "exact": dict(g1="exact", r="exact", g2w="exact", g3h="exact", g3r="exact"),
        "fp32-R": dict(g1="exact", r="exact", g2w="exact", g3h="exact", g3r="exact", fp32r=True),
        "all-tf32": dict(g1="tf32", r="tf32", g2w="tf32", g3h="tf32"),
        "g1-exact": dict(g1="exact", r="tf32", g2w="tf32", g3h="tf32"),
        "g1-exact+g2W-exact": dict(g1="exact", r="tf32", g2w="exact", g3h="tf32"),
        "g1-exact+g2W+g3H-exact": dict(g1="exact", r="tf32", g2w="exact", g3h="exact"),
        "only-g1-tf32": dict(g1="tf32", r="exact", g2w="exact", g3h="exact", g3r="exact"),
    }
    out = {}
    for name, mode in modes.items():
        W, H = W0.clone(), H0.clone()
        hist = []
        stop = None
        for it in range(1, iters + 1):
            W, H = step(X, W, H, mode)
            if it % 10 == 0:
                hist.append(kl(X, W, H))
                if stop is None and it >= 500 and len(hist) > 1 and abs(hist[-2] - hist[-1]) / abs(hist[-2]) < 1e-7:
                    stop = it
        out[name] = (W, hist, stop)
    Wx, hx, sx = out["exact"]
    print(f"--- D={D} k={k} iters={iters}: exact final KL {hx[-1]:.4f}, tol-1e-7 stop at {sx}")
    for name, (W, hist, stop) in out.items():
        cos = (W * Wx).sum(1) / (W.norm(dim=1) * Wx.norm(dim=1))
        print(f"{name:26s} KL rel gap {abs(hist[-1] - hx[-1]) / hx[-1]:.2e}  min cos {float(cos.min()):.7f}  stop {stop}")


if __name__ == "__main__":
    for D, k, iters in [(192, 8, 4000), (20000, 8, 3000), (50000, 20, 2000)]:
        run(D, k, iters)
