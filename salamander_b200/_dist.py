"""
Sample-sharded data parallelism: one process per GPU, contiguous row blocks
of ``adata.X``; H and every per-sample quantity stay local; only the V x k W numerator
(+ scalar objective, + k row sums for MvNMF) is summed across ranks each iteration.

These helpers are pure ``torch.distributed`` plumbing and work with any backend (NCCL on the
GPUs, gloo in the CPU tests of the host logic).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_samples: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, nearly equal row block [lo, hi) of rank ``rank``; the first ``n % world`` ranks get one extra."""
    base, extra = divmod(int(n_samples), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process).  NCCL's result is identical on all ranks."""
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_ranks_agree(flag: bool, device: torch.device) -> bool:
    """True iff ``flag`` is true on EVERY rank (all-reduce MIN): code paths whose kernels or collectives wait on the
    peers must be chosen by all ranks together."""
    if world()[1] == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather row blocks laid out by ``shard_bounds`` into the full (n_total, ...) tensor on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    counts = [shard_bounds(n_total, ws, r) for r in range(ws)]
    max_rows = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, counts)], dim=0)


def broadcast_numpy(arr: np.ndarray, device: torch.device, src: int = 0) -> np.ndarray:
    """Make rank ``src``'s host array the value on every rank (keeps replicas bit-identical after host init)."""
    rank, ws = world()
    if ws == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def exchange_owned_rows(t: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Every rank has written rows [lo, hi) of the replicated tensor ``t`` (its share of the signatures); make all rows
    current everywhere, in place.  Done as a sum of copies that are zero outside the owner's rows: x + 0 is exact, so the
    replicas stay bit-identical (multi-GPU CorrNMF signature embeddings)."""
    if world()[1] == 1:
        return t
    mine = torch.zeros_like(t)
    mine[lo:hi] = t[lo:hi]
    allreduce_sum_(mine)
    t.copy_(mine)
    return t


def allreduce_sum_counting_replicated_once(t: torch.Tensor, replicated: slice) -> torch.Tensor:
    """In-place sum over ranks of a vector whose entries in ``replicated`` are already identical on every rank (computed
    from replicated parameters) and must therefore be counted once, while the other entries are per-shard partial sums."""
    rank, ws = world()
    if ws == 1:
        return t
    if rank != 0:
        t[replicated] = 0
    return allreduce_sum_(t)


class PeerExchange:
    """Symmetric (peer-mapped) exchange buffer for the one-shot all-reduce fused into the reduction kernel
    (sal_klnmf_update_p2p): every rank allocates ``nbytes`` of zeroed symmetric memory, the ranks rendezvous and each
    keeps the table of all ranks' buffer addresses as mapped into ITS address space."""

    def __init__(self, nbytes: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        group = dist.group.WORLD
        self.buf = symm_mem.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.state = torch.tensor([1, 0], dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()  # every buffer is zeroed and mapped before anybody publishes into it


_PEER_EXCHANGES: dict[tuple[int, int, str], PeerExchange] = {}


def drop_peer_exchanges() -> None:
    """Forget the cached exchange buffers.  Called when a fit that used one ends with an exception: the ranks' tag counters may no
    longer agree, and words of the aborted update may still arrive.  The next fit allocates and zeroes a fresh buffer behind a
    barrier (every rank has to get there, which is what a collective abort means anyway)."""
    _PEER_EXCHANGES.clear()


def shared_peer_exchange(nbytes: int, device: torch.device, name: str = "klnmf") -> PeerExchange:
    """One exchange buffer per process, device and protocol (``name``), created collectively at first use and kept:
    allocating symmetric memory and the rendezvous cost ~100 ms, a fit at 8 GPUs lasts about as long.  Safe to share between
    consecutive fits: the sequence number that tags every word lives with the buffer and only ever increases, so words left
    over from an earlier fit (even one with another k, i.e. another layout) can never carry a tag a later update waits for.
    Every rank must ask for the same size in the same order (callers pass the size for the largest supported k)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), world()[1], name)
    px = _PEER_EXCHANGES.get(key)
    if px is None or px.nbytes < nbytes:
        px = PeerExchange(nbytes, device)
        px.nbytes = nbytes
        px.launch_id = 0
        _PEER_EXCHANGES[key] = px
    return px


LAUNCH_ID_LIMIT = 1 << 15


def next_launch_id(px: PeerExchange) -> int:
    """Launch number that prefixes the tags of a kernel exchanging through ``px`` (the CorrNMF signature-embedding solver): 1, 2,
    ... on every rank alike.  Before the 15-bit prefix would repeat, the buffers are zeroed behind a barrier and the count
    starts over -- every rank gets here at the same launch."""
    px.launch_id += 1
    if px.launch_id >= LAUNCH_ID_LIMIT:
        on_gpu = px.buf.is_cuda
        if on_gpu:
            torch.cuda.synchronize(px.buf.device)  # every kernel that reads or writes the buffers has finished here ...
        dist.barrier()                             # ... and on every peer
        px.buf.zero_()
        if on_gpu:
            torch.cuda.synchronize(px.buf.device)
        dist.barrier()
        px.launch_id = 1
    return px.launch_id
