"""
``KLNMF``: NMF under the (weighted) generalised KL divergence with normalised signatures and an
optional l-half sparsity penalty on the exposures.  Same interface as reference
models/klnmf.py:18-153; the numerics are the fused CUDA pass (sal_klnmf_pass) plus the
W epilogue (sal_w_epilogue).
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np
import torch

from .. import _dist
from .._device import PASS_OBJECTIVE, PASS_UPDATE_H, PASS_WNUM
from ..utils import shape_checker, type_checker
from .standard_nmf import StandardNMF

_FITTING_KWARGS = ["weights_kl", "weights_lhalf"]
_DEFAULT_FITTING_KWARGS = {kwarg: None for kwarg in _FITTING_KWARGS}


def period_launch_plan(min_iterations: int, max_iterations: int, freq: int, per_launch: int) -> list[tuple[int, int, bool]]:
    """The launches of the period driver if the convergence test never fires: ``(first period, number of periods, trailing
    objective)``.  Periods whose incoming checkpoint cannot end the fit (p = 0 or p * freq < min_iterations) are folded, up to
    ``per_launch`` of them, into one launch; every later period is its own (speculative) launch; the launch that ends at
    ``max_iterations`` carries the objective of the final iterate if that is a checkpoint (reference signature_nmf.py:365-380).
    A pure function of its arguments: every rank of a sharded fit issues the same sequence."""
    freq, max_it = int(freq), max(int(max_iterations), 1)
    n_seg, final_ckpt = -(-max_it // freq), max_it % freq == 0

    def deciding(p):
        return p >= 1 and p * freq >= int(min_iterations)

    plan, seg = [], 0
    while seg < n_seg:
        m = 1
        if not deciding(seg):
            while seg + m < n_seg and not deciding(seg + m) and m < per_launch:
                m += 1
        plan.append((seg, m, bool(final_ckpt and seg + m == n_seg)))
        seg += m
    return plan


class KLNMF(StandardNMF):
    def __init__(
        self,
        n_signatures: int = 1,
        init_method: str = "nndsvd",
        min_iterations: int = 500,
        max_iterations: int = 10000,
        conv_test_freq: int = 10,
        tol: float = 1e-7,
        **device_kwargs,
    ):
        super().__init__(n_signatures, init_method, min_iterations, max_iterations, conv_test_freq, tol, **device_kwargs)
        self.weights_kl = None
        self.weights_lhalf = None
        # capture the periods of the fit loop in CUDA graphs (see _fit_loop): True, False or "auto" = only when a shard is
        # small enough for launch gaps and host time to matter (_GRAPH_MAX_SAMPLES).  Measured on B200 (profiles/r01_e2e_phases.md): at 1M samples an eager
        # period is within 2 % of a replayed one while capturing ~6 graphs per fit costs more than that and occasionally
        # stalls in the driver; at 125k samples graphs are 12 % faster.
        self.use_graphs: bool | str = "auto"
        # tf32 fits: a whole convergence-test period per persistent launch, reduction / exchange / W epilogue inside the
        # kernel (sal_klnmf_period, see _fit_loop_period); False falls back to the two-kernel updates below
        self.use_period_kernel = True
        self.loop_start_event = None  # optional torch.cuda.Event the period driver records right before its first launch
        self.loop_end_event = None    # ... and right behind the last planned launch and the copy of its objectives to the host
        self.objectives_to_host = True  # period kernel: objectives stored straight into page-locked host memory
        self.use_small_kernel = True  # problems that fit one SM: persistent single-CTA kernel (see _fit_loop_small)
        # multi-GPU all-reduce of the numerator: "auto" / "p2p" = one-shot NVLink exchange fused into the reduction
        # kernel (sal_klnmf_update_p2p), "nccl" = library collective between separate kernels
        self.allreduce = "auto"

    @property
    def objective(self) -> Literal["minimize", "maximize"]:
        return "minimize"

    def _upload_fitting_parameters(self) -> None:
        st = self._dev
        st.weights["kl"] = st.shard(self.weights_kl)
        st.weights["lhalf"] = st.shard(self.weights_lhalf)

    def objective_function(self) -> float:
        """(Weighted) KL divergence plus the l-half penalty (reference klnmf.py:64-80)."""
        with self._resident() as st:
            st.ws.klnmf_pass(
                st.X, st.W, st.H, PASS_OBJECTIVE, w_kl=st.weights["kl"], w_lhalf=st.weights["lhalf"], objective=st.obj
            )
            return st.objective_value()

    def _update_parameters(self, given_parameters: dict[str, Any] | None = None) -> None:
        """One joint multiplicative update of W and H (reference klnmf.py:86-106 -> update_WH).

        Single GPU: two launches (sal_klnmf_update: fused pass, then reduction + W epilogue).  Several GPUs: the pass
        leaves this rank's numerator, the 96 x k numerators are summed over the ranks and every rank applies the
        same epilogue."""
        n_given = self._n_given(given_parameters)
        with self._resident() as st:
            if st.world == 1:
                st.ws.klnmf_update(
                    st.X, st.W, st.W_next, st.H, st.H, n_given, True, st.Wnum, w_kl=st.weights["kl"], w_lhalf=st.weights["lhalf"]
                )
                st.W, st.W_next = st.W_next, st.W
                return
            flags = PASS_UPDATE_H | (PASS_WNUM if n_given < st.k else 0)
            st.ws.klnmf_pass(
                st.X,
                st.W,
                st.H,
                flags,
                H_out=st.H,
                w_kl=st.weights["kl"],
                w_lhalf=st.weights["lhalf"],
                Wnum=st.Wnum,
            )
            if n_given < st.k:
                st.allreduce(st.Wnum)
                st.ws.w_epilogue(st.W, st.Wnum, n_given, True, st.W)

    # ---- device-side fit driver ------------------------------------------------------------------
    def _one_update(self, st, n_given, W_in, W_out, H_in, H_out, objective=None) -> None:
        """Joint update (W_in, H_in) -> (W_out, H_out); ``objective`` (device scalar) receives the objective of
        the INCOMING iterate.  Single GPU: two launches.  Several GPUs: pass, reduction, all-reduce of the 96 x k
        numerator (and of the objective scalar), W epilogue."""
        wk, wl = st.weights["kl"], st.weights["lhalf"]
        if st.world == 1:
            st.ws.klnmf_update(st.X, W_in, W_out, H_in, H_out, n_given, True, st.Wnum, w_kl=wk, w_lhalf=wl, objective=objective)
            return
        px = self._peer_exchange(st) if n_given < st.k else None
        if px is not None:
            st.ws.klnmf_update_p2p(
                st.X, W_in, W_out, H_in, H_out, n_given, True, st.Wnum, px.peers, px.state, px.world, px.rank,
                w_kl=wk, w_lhalf=wl, objective=objective,
            )
            return
        flags = PASS_UPDATE_H | (PASS_WNUM if n_given < st.k else 0) | (PASS_OBJECTIVE if objective is not None else 0)
        st.ws.klnmf_pass(st.X, W_in, H_in, flags, H_out=H_out, w_kl=wk, w_lhalf=wl, Wnum=st.Wnum, objective=objective)
        if n_given < st.k:
            st.allreduce(st.Wnum)
        if objective is not None:
            st.allreduce(objective)
        st.ws.w_epilogue(W_in, st.Wnum, n_given, True, W_out)

    def _fit_loop_small(self, n_given: int, verbose, verbosity_freq):
        """Fit driver for problems that fit one SM (sal_klnmf_small_updates; BASELINE config 0): one launch of a
        persistent single-CTA kernel per period -- the objective of the period's incoming iterate plus its
        ``conv_test_freq`` updates -- and the NEXT period is launched before the host looks at that objective, so the GPU
        never waits for Python.  States rotate through three buffers: the one the convergence test may still fall back
        to is never overwritten.  Same iterates, history and stopping iteration as the reference loop."""
        st = self._dev
        freq, max_it, min_it = int(self.conv_test_freq), int(self.max_iterations), int(self.min_iterations)
        Wb = [st.W, torch.empty_like(st.W), torch.empty_like(st.W)]
        Hb = [st.H, torch.empty_like(st.H), torch.empty_like(st.H)]
        obj_dev = torch.zeros(2, dtype=torch.float64, device=st.device)
        obj_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        events = [torch.cuda.Event(), torch.cuda.Event()]

        def n_of(i: int) -> int:
            return min(i * freq, max_it)

        def launch(i: int) -> None:  # state i -> state i + 1, objective of state i
            n_i, s = n_of(i), i % 2
            st.ws.klnmf_small_updates(
                st.X, Wb[i % 3], Wb[(i + 1) % 3], Hb[i % 3], Hb[(i + 1) % 3], n_given, min(freq, max_it - n_i), objective=obj_dev[s : s + 1]
            )
            obj_host[s : s + 1].copy_(obj_dev[s : s + 1], non_blocking=True)
            events[s].record()

        of_values: list[float] = []
        i = 0
        launch(0)
        while True:
            n_i, n_next = n_of(i), n_of(i + 1)
            speculative = n_i < max_it and n_next % freq == 0
            if speculative:
                launch(i + 1)
            events[i % 2].synchronize()
            of_values.append(float(obj_host[i % 2]))
            final = i
            if i > 0:
                rel_change = np.abs(of_values[-2] - of_values[-1]) / np.abs(of_values[-2])
                if bool(rel_change < self.tol and n_i >= min_it):
                    break  # the fit ended at iteration n_i; later periods are dropped
            if n_i >= max_it:
                break
            for it in range(n_i + 1, n_next + 1):
                if verbose and it % verbosity_freq == 0:
                    print(f"iteration: {it}; objective: {of_values[-1]:.2f}")
            final = i + 1
            if not speculative:  # the last, shorter period ends at max_iterations without an objective
                break
            i += 1
        torch.cuda.current_stream(st.device).synchronize()
        st.W, st.H = Wb[final % 3], Hb[final % 3]
        st.W_next = Wb[(final + 1) % 3]
        self.launch_stats = {"graphs": 0, "periods": i + 1, "driver": "single-CTA persistent kernel"}
        return of_values, n_of(final)

    def _peer_exchange(self, st):
        """The symmetric exchange buffer of the fused all-reduce, set up once per device state; ``None`` (NCCL path)
        when ``allreduce = "nccl"`` was asked for or symmetric memory cannot be set up on EVERY rank: the choice is
        made collectively (a rank that fell back alone would leave its peers polling for words that never arrive)."""
        if self.allreduce == "nccl":
            return None
        if "peer_exchange" not in st.weights:
            px, err = None, None
            try:
                nbytes = int(st.ws.lib.sal_p2p_exchange_bytes(32, st.world))  # sized for the largest k: shared by all fits
                px = _dist.shared_peer_exchange(nbytes, st.device)
            except Exception as exc:  # pragma: no cover - depends on the system
                err = exc
            # (a rank with an empty shard cannot take part in the exchange kernel either: it has no partials to publish)
            if not _dist.all_ranks_agree(px is not None and (st.hi - st.lo) > 0, st.device):
                if self.allreduce == "p2p":
                    raise RuntimeError(f"peer-memory all-reduce unavailable on at least one rank ({err})")
                import warnings

                warnings.warn(f"peer-memory all-reduce unavailable on at least one rank ({err}); using NCCL")
                px = None
            st.weights["peer_exchange"] = px
        return st.weights["peer_exchange"]

    # ---- period-kernel fit driver ------------------------------------------------------------------
    _PERIODS_PER_LAUNCH = 16  # non-deciding periods folded into one launch
    _QUEUE_DEPTH = 4          # launches in flight while no decision is pending

    def _period_path(self, st, n_given):
        """(use the persistent period kernel?, peer exchange).  Decided collectively on several GPUs, once per device state
        (the agreement is a collective plus a host synchronisation: not something to repeat per fit loop)."""
        if not self.use_period_kernel or st.weights["kl"] is not None or st.weights["lhalf"] is not None or st.ws.timing:
            return False, None
        key = ("period_path", n_given)
        if key not in st.weights:
            px = None
            ok = st.ws.period_supported(n_given, st.world)
            if st.world > 1:
                px = self._peer_exchange(st) if ok else None
                ok = _dist.all_ranks_agree(ok and px is not None, st.device)
            st.weights[key] = (ok, px)
        return st.weights[key]

    def _fit_loop_period(self, st, px, n_given, verbose, verbosity_freq):
        """Fit driver on the persistent period kernel (sal_klnmf_period).  Same iterates, history and stopping iteration as
        the reference loop (signature_nmf.py:361-380):

        * a launch runs whole periods of ``conv_test_freq`` updates, the objective of each period's incoming iterate fused
          into its first update and -- when the launch ends at ``max_iterations`` -- the objective of the final iterate as a
          trailing sweep; nothing else touches X;
        * while the convergence test cannot fire (n < ``min_iterations``) there is nothing to decide: several periods go
          into one launch and the host runs ahead of the GPU;
        * afterwards every launch is one period and speculative: it writes into spare buffers (three W / H states rotate)
          and up to two periods are in flight beyond the last objective the host has seen; a period whose incoming
          iterate turns out to be the converged one is simply not adopted.
        """
        freq, min_it = int(self.conv_test_freq), int(self.min_iterations)
        max_it = max(int(self.max_iterations), 1)
        fl = st.fit_loop
        if not fl.get("period") or fl["Ws"][fl["cur"]] is not st.W or fl["Hs"][fl["cur"]] is not st.H:
            fl.clear()
            fl.update({"period": True, "cur": 0, "Ws": [st.W, st.W_next, torch.empty_like(st.W)],
                       "Hs": [st.H, torch.empty_like(st.H), torch.empty_like(st.H)]})
        Ws, Hs, cur = fl["Ws"], fl["Hs"], fl["cur"]
        n_slots = self._PERIODS_PER_LAUNCH + 1
        ring = 2 * self._QUEUE_DEPTH
        if "obj_dev" not in fl:
            fl["obj_dev"] = torch.zeros((ring, n_slots), dtype=torch.float64, device=st.device)
            fl["obj_host"] = torch.zeros((ring, n_slots), dtype=torch.float64).pin_memory()
            fl["events"] = [torch.cuda.Event() for _ in range(ring)]
        obj_dev, obj_host, events = fl["obj_dev"], fl["obj_host"], fl["events"]

        def deciding(p):  # can the convergence test fire at checkpoint p (iteration p * freq)?
            return p >= 1 and p * freq >= min_it

        pending = []  # (ring slot, first checkpoint, n checkpoints, state index before, state index after)
        plan = period_launch_plan(min_it, max_it, freq, self._PERIODS_PER_LAUNCH)
        nxt_launch, n_launched, of_values, n_done, final_state = 0, 0, [], 0, None
        while True:
            # ---- launch as far ahead as the decisions allow ----
            while nxt_launch < len(plan):
                seg, m, fin = plan[nxt_launch]
                if deciding(seg):
                    if sum(1 for q in pending if deciding(q[1])) >= 2:
                        break
                elif len(pending) >= self._QUEUE_DEPTH:
                    break
                n0, n1 = seg * freq, min((seg + m) * freq, max_it)
                slot, nxt = n_launched % ring, (cur + 1) % 3
                if n_launched == 0 and self.loop_start_event is not None:
                    self.loop_start_event.record()  # measurement aid: device time from the driver's first launch on
                # (the kernel writes its objectives straight into the page-locked ring: no copy behind the launch)
                st.ws.klnmf_period(
                    st.X, Ws[cur], Ws[nxt], Hs[cur], Hs[nxt], n_given, True, n1 - n0, freq, fin,
                    objectives=obj_host[slot] if self.objectives_to_host else obj_dev[slot],
                    peers=None if px is None else px.peers, state=None if px is None else px.state,
                    n_ranks=st.world if px is not None else 1, rank=st.rank if px is not None else 0,
                )
                if not self.objectives_to_host:
                    obj_host[slot].copy_(obj_dev[slot], non_blocking=True)
                events[slot].record()
                if nxt_launch + 1 == len(plan) and self.loop_end_event is not None:
                    self.loop_end_event.record()  # measurement aid: everything the loop asked of the device has been enqueued
                pending.append((slot, seg, m + (1 if fin else 0), cur, nxt))
                cur, nxt_launch, n_launched = nxt, nxt_launch + 1, n_launched + 1
            if not pending:
                break
            # ---- the oldest launch's objectives ----
            slot, p0, n_ck, s_before, s_after = pending.pop(0)
            events[slot].synchronize()
            stop = False
            for i in range(n_ck):
                p = p0 + i
                of_values.append(float(obj_host[slot, i]))
                if p >= 1:
                    rel_change = np.abs(of_values[-2] - of_values[-1]) / np.abs(of_values[-2])
                    if bool(rel_change < self.tol and p * freq >= min_it):
                        # the fit ended at iteration p * freq.  A deciding launch holds one period, so that is the state it
                        # started from (it and the launches behind it are not adopted) -- or, for the trailing checkpoint at
                        # max_iterations, the state it ended with
                        n_done, final_state, stop = p * freq, (s_before if i == 0 else s_after), True
                        break
                if p * freq >= max_it:
                    n_done, final_state, stop = max_it, s_after, True
                    break
                n_end = min((p + 1) * freq, max_it)
                for it in range(p * freq + 1, n_end + 1):
                    if verbose and it % verbosity_freq == 0:
                        print(f"iteration: {it}; objective: {of_values[-1]:.2f}")
                n_done, final_state = n_end, s_after
            if stop or (nxt_launch >= len(plan) and not pending):
                break
        torch.cuda.current_stream(st.device).synchronize()  # dropped speculative launches still write the spare buffers
        st.W, st.H = Ws[final_state], Hs[final_state]
        st.W_next = Ws[(final_state + 1) % 3]
        fl["cur"] = final_state
        self.launch_stats = {"graphs": 0, "launches": n_launched, "driver": "persistent period kernel"}
        return of_values, n_done

    # "auto": graphs while an update takes less than ~80 us of GPU time.  An eager update costs the host ~50-60 us (ctypes
    # marshalling, tensor-map encodes, two launches): hidden behind the GPU at 1M samples per shard, not at 500k (measured
    # at 2 GPUs: 65.8 us per update eager vs 60.8 us replayed)
    _GRAPH_MAX_SAMPLES = 800_000

    def _fit_loop(self, given_parameters, verbose, verbosity_freq):
        """Same iterates, history and stopping iteration as the reference loop (signature_nmf.py:361-380), run in
        periods of ``conv_test_freq`` updates:

        * the objective of iteration n is produced by the fused pass of update n + 1 (it streams X anyway), not by a
          separate pass over X;
        * the period's updates are speculative: they write into spare buffers, state n stays untouched until the
          host has seen the objective, and the period is dropped if the convergence test says the fit ended at n;
        * each period is two CUDA graphs (update n + 1 with the objective read-back | the remaining updates), so the
          host issues two launches per period and the GPU never waits for Python.
        """
        st = self._dev
        freq, max_it = int(self.conv_test_freq), int(self.max_iterations)
        n_given = self._n_given(given_parameters)
        if (
            self.use_small_kernel
            and st.world == 1
            and st.weights["kl"] is None
            and st.weights["lhalf"] is None
            and not st.ws.timing
            and st.ws.small_supported()
        ):
            return self._fit_loop_small(n_given, verbose, verbosity_freq)
        use_period, px = self._period_path(st, n_given)
        if use_period:
            try:
                return self._fit_loop_period(st, px, n_given, verbose, verbosity_freq)
            except BaseException:
                if px is not None:  # the exchange protocol may be out of step now: never reuse this buffer
                    _dist.drop_peer_exchanges()
                    st.weights.pop("peer_exchange", None)
                raise
        if freq < 3:
            return super()._fit_loop(given_parameters, verbose, verbosity_freq)
        # spare buffers, pinned read-back slot and captured graphs live with the device state, so that a second
        # fit loop on the same state (bench.py: warm-up, then the timed run) replays instead of re-capturing
        fl = st.fit_loop
        if not fl or not any(w is st.W for w in fl["Ws"]) or not any(h is st.H for h in fl["Hs"]) or fl["n_given"] != n_given:
            fl = st.fit_loop = {
                "Ws": [st.W, st.W_next, torch.empty_like(st.W)],
                "Hs": [st.H, torch.empty_like(st.H)],
                "obj_host": torch.zeros(1, dtype=torch.float64).pin_memory(),
                "graphs": {},
                "periods": 0,
                "n_given": n_given,
            }
        Ws, Hs, obj_host, graphs = fl["Ws"], fl["Hs"], fl["obj_host"], fl["graphs"]
        seen = torch.cuda.Event()
        use_graphs = self.use_graphs
        if use_graphs == "auto":
            use_graphs = (st.hi - st.lo) <= self._GRAPH_MAX_SAMPLES
        use_graphs = bool(use_graphs) and not st.ws.timing

        def plan(wi, hi, L):
            """Buffer indices of the L updates of a period that starts from (Ws[wi], Hs[hi])."""
            others = [j for j in range(3) if j != wi]
            steps, w_prev, h_other = [], wi, 1 - hi
            for u in range(L):
                w_out = others[u % 2]
                steps.append((w_prev, w_out, hi if u == 0 else h_other, h_other))
                w_prev = w_out
            return steps, w_prev, h_other

        def run_head(step):
            self._one_update(st, n_given, Ws[step[0]], Ws[step[1]], Hs[step[2]], Hs[step[3]], objective=st.obj)
            obj_host.copy_(st.obj, non_blocking=True)

        def run_tail(steps):
            for step in steps:
                self._one_update(st, n_given, Ws[step[0]], Ws[step[1]], Hs[step[2]], Hs[step[3]])

        def launch(wi, hi, L, periods_done):
            steps, w_end, h_end = plan(wi, hi, L)
            key = (wi, hi, L)
            if use_graphs and periods_done >= 2 and key not in graphs:
                # every kernel variant and the NCCL communicator have been exercised by the eager periods
                # (raw capture_begin / capture_end on a side stream: torch.cuda.graph() would synchronise the device and
                # empty the device and pinned-host allocator caches at every capture)
                g_head, g_tail = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                main = torch.cuda.current_stream(st.device)
                side = fl.setdefault("capture_stream", torch.cuda.Stream(st.device))
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for g, body, arg in ((g_head, run_head, steps[0]), (g_tail, run_tail, steps[1:])):
                        g.capture_begin(capture_error_mode="thread_local")
                        try:
                            body(arg)
                        finally:
                            g.capture_end()
                main.wait_stream(side)
                graphs[key] = (g_head, g_tail)
            if use_graphs and key in graphs:
                graphs[key][0].replay()
                seen.record()
                graphs[key][1].replay()
            else:
                run_head(steps[0])
                seen.record()
                run_tail(steps[1:])
            return w_end, h_end

        of_values: list[float] = []
        n, periods = 0, fl["periods"]
        wi = next(j for j, w in enumerate(Ws) if w is st.W)
        hi = next(j for j, h in enumerate(Hs) if h is st.H)
        while True:
            L = min(freq, max_it - n)
            if L <= 0:  # n == max_iterations, a multiple of conv_test_freq: the reference still evaluates the objective
                st.W, st.H = Ws[wi], Hs[hi]
                of_values.append(self.objective_function())
                break
            w_end, h_end = launch(wi, hi, L, periods)
            periods += 1
            seen.synchronize()
            of_values.append(float(obj_host.item()))
            if n > 0:
                rel_change = np.abs(of_values[-2] - of_values[-1]) / np.abs(of_values[-2])
                if bool(rel_change < self.tol and n >= self.min_iterations):
                    break  # the fit ended at iteration n; the speculative period is not adopted
            wi, hi = w_end, h_end
            for it in range(n + 1, n + L + 1):
                if verbose and it % verbosity_freq == 0:
                    print(f"iteration: {it}; objective: {of_values[-1]:.2f}")
            n += L
            if n >= max_it and n % freq != 0:
                break
        torch.cuda.current_stream(st.device).synchronize()
        st.W, st.H = Ws[wi], Hs[hi]
        st.W_next = Ws[(wi + 1) % 3]
        fl["periods"] = periods
        self.launch_stats = {"graphs": 2 * len(graphs), "periods": periods}
        return of_values, n

    def _check_weights(self, weights: np.ndarray, name: str = "weights") -> None:
        """Type, shape and sign of per-sample weights (reference klnmf.py:108-126)."""
        type_checker(name, weights, np.ndarray)
        shape_checker(name, weights, (self.adata.n_obs,))
        if not all(weights >= 0):
            raise ValueError("Only non-negative KL-divergence and sparsity penalty weights are allowed.")

    def _setup_fitting_parameters(self, fitting_kwargs: dict[str, Any] | None = None) -> None:
        """Scalars / lists are broadcast to (n_obs,) arrays (reference klnmf.py:128-153)."""
        if fitting_kwargs is None:
            fitting_kwargs = _DEFAULT_FITTING_KWARGS
        for kwarg in fitting_kwargs:
            if kwarg not in _FITTING_KWARGS:
                raise ValueError(
                    f"The given fitting keyword arguments include parameters outside of {_FITTING_KWARGS}."
                )
        for name, weights in fitting_kwargs.items():
            if weights is not None:
                type_checker(name, weights, [float, int, list, np.ndarray])
                if type(weights) in [float, int]:
                    weights = weights * np.ones(self.adata.n_obs)
                if type(weights) is list:
                    weights = np.array(weights)
                self._check_weights(weights, name)
            setattr(self, name, weights)
