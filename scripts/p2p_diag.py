"""Diagnostics (2+ GPUs, torchrun): per-iteration cost of the fused reduce + NVLink all-reduce + epilogue kernel vs NCCL."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
import salamander_b200 as sal
from salamander_b200 import AnnData, _dist
from salamander_b200._device import PASS_UPDATE_H, PASS_WNUM

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev); world = dist.get_world_size()
D, k = 1_000_000, 20
lo, hi = bench.shard_bounds(D, world, rank)
X = bench.synth_rows(lo, hi, k); W0, H0 = bench.init_rows(X, lo, k)
m = sal.models.KLNMF(n_signatures=k, init_method="custom", dtype="float32", math="tf32", device=dev, shard_input=False)
m._setup_adata(AnnData(X)); m._initialize(None, {"signatures_mat": W0, "exposures_mat": H0}); m._setup_fitting_parameters(None); m._to_device()
st = m._dev
W2 = torch.empty_like(st.W)
px = m._peer_exchange(st)

def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def pass_only(): st.ws.klnmf_pass(st.X, st.W, st.H, PASS_UPDATE_H | PASS_WNUM, H_out=st.H, Wnum=st.Wnum)
def p2p(): st.ws.klnmf_update_p2p(st.X, st.W, W2, st.H, st.H, 0, True, st.Wnum, px.peers, px.state, px.world, px.rank)
def nccl():
    st.ws.klnmf_pass(st.X, st.W, st.H, PASS_UPDATE_H | PASS_WNUM, H_out=st.H, Wnum=st.Wnum)
    dist.all_reduce(st.Wnum); st.ws.w_epilogue(st.W, st.Wnum, 0, True, W2)
def graphed(fn, n=10):
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    return lambda: g.replay()
res = {"pass+finish": timeit(pass_only), "p2p eager": timeit(p2p), "nccl eager": timeit(nccl)}
gp = graphed(p2p); res["p2p graph(10)"] = timeit(gp, 50) / 10
gn = graphed(nccl); res["nccl graph(10)"] = timeit(gn, 50) / 10
if rank == 0: print({k_: round(v, 1) for k_, v in res.items()}, "us per iteration, world", world)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
