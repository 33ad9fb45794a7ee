"""
``MvNMF``: minimum-volume NMF, KL divergence + lam * ln det(W^T W + delta I), with the
reference's H step, closed-form unconstrained W step and back-tracking line search
(reference models/mvnmf.py:95-218).  The single-step methods (``_update_H``, ``_update_W``, ``_line_search``; the reference's
tests call them) do, per iteration:

    pass 1  UPDATE_H                          (update_H,               _utils_klnmf.py:220-264)
    pass 2  WNUM | HSUM | OBJECTIVE on new H  (N = (X/WH) H^T, rowsums, previous objective)
    1 CTA   logdet(W), W_unconstrained        (mvnmf.py:19-24, 37-66)
    per line-search trial: 1 CTA blend/normalise/clip/logdet + pass OBJECTIVE|UPDATE_H with
            h_scale, writing the candidate H into a spare buffer that is swapped in on accept.

``fit`` runs the same arithmetic through one of three drivers: the persistent single-CTA / thread-block-cluster kernel for
problems that fit shared memory (whole iterations incl. the line search per launch), or -- larger problems -- the run-ahead
driver: TWO passes over X per iteration (the next H step rides on the accepted trial's pass, SAL_PASS_SCALED_UPDATE), the
unconstrained step and the first candidate in one launch, and an optimistic line search whose decision the host reads one
iteration later (roll-back when the full step was rejected).  The reference needs >= 4 passes over X per iteration plus one
per back-track.
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np
import torch

from .. import _dist
from .._device import PASS_HSUM, PASS_OBJECTIVE, PASS_SCALED_UPDATE, PASS_UPDATE_H, PASS_WNUM
from .standard_nmf import StandardNMF


class MvNMF(StandardNMF):
    def __init__(
        self,
        n_signatures: int = 1,
        init_method: str = "nndsvd",
        lam: float = 1.0,
        delta: float = 1.0,
        min_iterations: int = 500,
        max_iterations: int = 10000,
        conv_test_freq: int = 10,
        tol: float = 1e-7,
        **device_kwargs,
    ):
        super().__init__(n_signatures, init_method, min_iterations, max_iterations, conv_test_freq, tol, **device_kwargs)
        self.lam = lam
        self.delta = delta
        self._gamma = 1.0
        self.use_small_kernel = True  # problems that fit one SM: whole iterations in a persistent single-CTA kernel
        self.run_ahead = True         # larger problems: optimistic line search, the host runs ahead of the GPU
        self.fuse_h_step = True       # ... and the next iteration's H step rides on the line-search trial's pass: 2 passes per iteration

    @property
    def objective(self) -> Literal["minimize", "maximize"]:
        return "minimize"

    def _upload_fitting_parameters(self) -> None:
        st = self._dev
        st.weights["kl"] = None
        st.weights["lhalf"] = None
        st.W_unc = torch.empty_like(st.W)
        st.W_unc_spare = torch.empty_like(st.W)  # run-ahead driver: the previous iteration's W_unconstrained stays intact
        st.W_spare = torch.empty_like(st.W)      # ... and so does its W while the next candidate is already being written
        st.W_trial = torch.empty_like(st.W)
        st.H_trial = torch.empty_like(st.H)
        st.h_scale = torch.empty(st.k, dtype=st.dtype, device=st.device)
        st.hsum = torch.empty(st.k, dtype=st.dtype, device=st.device)
        st.kl2 = torch.zeros(2, dtype=torch.float64, device=st.device)  # [previous, trial] KL (summed over ranks)
        st.ld2 = torch.zeros(2, dtype=torch.float64, device=st.device)  # [previous, trial] logdet (replicated)

    def objective_function(self) -> float:
        """KL + lam * logdet volume (reference mvnmf.py:149-156)."""
        with self._resident() as st:
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_OBJECTIVE, objective=st.kl2[0:1])
            st.ws.mvnmf_logdet(st.W, self.delta, st.ld2[0:1])
            st.allreduce(st.kl2[0:1])
            kl, ld = torch.stack([st.kl2[0], st.ld2[0]]).tolist()
            return kl + self.lam * ld

    def _update_H(self) -> None:
        with self._resident() as st:
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_UPDATE_H, H_out=st.H)

    def _update_W_unconstrained(self, n_given_signatures: int = 0, logdet_known: bool = False, with_first_trial: bool = False) -> None:
        """Leaves W_unconstrained in ``st.W_unc`` and the previous objective parts in kl2[0], ld2[0].  ``logdet_known``: W is the
        candidate the last line-search trial accepted, whose log-determinant that trial left in ld2[1] (same kernel
        arithmetic on the same numbers): copied instead of recomputed."""
        st = self._dev
        st.ws.klnmf_pass(
            st.X, st.W, st.H, PASS_WNUM | PASS_HSUM | PASS_OBJECTIVE, Wnum=st.Wnum, hsum=st.hsum, objective=st.kl2[0:1]
        )
        st.allreduce(st.Wnum)
        st.allreduce(st.hsum)
        if logdet_known:
            st.ld2[0:1].copy_(st.ld2[1:2])
        else:
            st.ws.mvnmf_logdet(st.W, self.delta, st.ld2[0:1])
        if with_first_trial:  # ... and the first line-search candidate (the full step) in the same launch
            st.ws.mvnmf_w_unconstrained_trial(st.W, st.Wnum, st.hsum, self.lam, self.delta, n_given_signatures, st.W_unc,
                                              st.W_trial, st.h_scale, st.ld2[1:2])
        else:
            st.ws.mvnmf_w_unconstrained(st.W, st.Wnum, st.hsum, self.lam, self.delta, n_given_signatures, st.W_unc)

    def _trial(self, gamma_blend: float, fuse_h_step: bool = False, candidate_ready: bool = False) -> float:
        """One line-search candidate: W_trial, the rescaled exposures and the candidate's objective parts (kl2[1], ld2[1]).
        ``fuse_h_step``: the pass also applies the NEXT iteration's H step to the rescaled exposures (it has W_trial, the
        rescaled exposures and the quotient in hand: SAL_PASS_SCALED_UPDATE) -- H_trial then holds update_H's result and the
        next iteration starts at its W step."""
        st = self._dev
        if not candidate_ready:
            st.ws.mvnmf_trial(st.W, st.W_unc, gamma_blend, self.delta, st.W_trial, st.h_scale, st.ld2[1:2])
        st.ws.klnmf_pass(
            st.X,
            st.W_trial,
            st.H,
            PASS_OBJECTIVE | PASS_UPDATE_H | (PASS_SCALED_UPDATE if fuse_h_step else 0),
            H_out=st.H_trial,
            h_scale=st.h_scale,
            objective=st.kl2[1:2],
        )

    def _line_search(self) -> None:
        """Back-tracking on the penalised objective; the first trial ignores gamma (reference mvnmf.py:69-92)."""
        st = self._dev
        self._trial(-1.0)
        st.allreduce(st.kl2)
        kl_prev, kl_new, ld_prev, ld_new = torch.cat([st.kl2, st.ld2]).tolist()
        prev_of_value = kl_prev + self.lam * ld_prev
        of_value = kl_new + self.lam * ld_new
        gamma = self._gamma
        while of_value > prev_of_value and gamma > 1e-16:
            gamma *= 0.8
            self._trial(gamma)
            st.allreduce(st.kl2[1:2])
            kl_new, ld_new = torch.stack([st.kl2[1], st.ld2[1]]).tolist()
            of_value = kl_new + self.lam * ld_new
        self._gamma = min(1.0, 1.2 * gamma)
        st.W, st.W_trial = st.W_trial, st.W
        st.H, st.H_trial = st.H_trial, st.H

    def _update_W(self, n_given_signatures: int = 0) -> None:
        if n_given_signatures == self.n_signatures:
            return
        with self._resident():
            self._update_W_unconstrained(n_given_signatures)
            self._line_search()

    def _update_parameters(self, given_parameters: dict[str, Any] | None = None) -> None:
        """H step, then the W step unless all signatures are given (reference mvnmf.py:197-210)."""
        with self._resident():
            self._update_H()
            self._update_W(self._n_given(given_parameters))

    # ---- fit drivers -------------------------------------------------------------------------------------
    def _fit_loop(self, given_parameters, verbose, verbosity_freq):
        st = self._dev
        n_given = self._n_given(given_parameters)
        if self.use_small_kernel and st.world == 1 and not st.ws.timing and st.ws.mvnmf_small_supported():
            return self._fit_loop_small(n_given, verbose, verbosity_freq)
        if self.run_ahead and n_given < self.n_signatures and not st.ws.timing:
            return self._fit_loop_run_ahead(n_given, verbose, verbosity_freq)
        return super()._fit_loop(given_parameters, verbose, verbosity_freq)

    # -- optimistic line search: the host never waits for the iteration it has just launched ---------------
    # _first_trial launches the full step (reference mvnmf.py:80-88) and adopts the candidate without looking at the objective;
    # the four numbers the decision needs travel to a pinned slot behind it, and _confirm checks them later.
    def _first_trial(self, ring, fused: bool) -> dict:
        st = self._dev
        self._trial(-1.0, fused, candidate_ready=True)  # (the candidate came out of _update_W_unconstrained's launch)
        st.allreduce(st.kl2)
        slot = ring["next"] % len(ring["events"])
        ring["next"] += 1
        ring["host"][slot].copy_(torch.cat([st.kl2, st.ld2]), non_blocking=True)
        ring["events"][slot].record()
        rec = {"slot": slot, "W": st.W, "W_trial": st.W_trial, "W_spare": st.W_spare, "H": st.H, "H_trial": st.H_trial,
               "W_unc": st.W_unc, "W_unc_spare": st.W_unc_spare, "gamma": self._gamma, "fused": fused}
        # adopt the candidate.  The next iteration's unconstrained step AND first candidate are launched before this iteration
        # has been confirmed: they must overwrite neither this iteration's W (a back-track blends with it) nor its
        # W_unconstrained -- three W buffers rotate (current, next candidate, the one kept), two W_unconstrained buffers
        st.W, st.W_trial, st.W_spare = st.W_trial, st.W_spare, st.W
        st.H, st.H_trial = st.H_trial, st.H
        st.W_unc, st.W_unc_spare = st.W_unc_spare, st.W_unc
        self._gamma = min(1.0, 1.2 * self._gamma)
        return rec

    def _confirm(self, rec, ring) -> bool:
        """Was the optimistic iteration's full step accepted?  If not: drop whatever was launched on top of it, restore the
        iteration's buffers (nothing that a back-track reads has been overwritten: the driver runs at most the H step and the
        unconstrained W step of the NEXT iteration ahead, which touch neither the kept exposures, the previous W nor the kept
        W_unconstrained) and run the reference's back-tracking loop.  Returns False when it had to do that."""
        st = self._dev
        ring["events"][rec["slot"]].synchronize()
        kl_prev, kl_new, ld_prev, ld_new = ring["host"][rec["slot"]].tolist()
        prev_of_value = kl_prev + self.lam * ld_prev
        of_value = kl_new + self.lam * ld_new
        if not of_value > prev_of_value:
            return True
        torch.cuda.current_stream(st.device).synchronize()
        st.W, st.W_trial, st.W_spare, st.H, st.H_trial = rec["W"], rec["W_trial"], rec["W_spare"], rec["H"], rec["H_trial"]
        st.W_unc, st.W_unc_spare = rec["W_unc"], rec["W_unc_spare"]
        gamma = rec["gamma"]
        while of_value > prev_of_value and gamma > 1e-16:
            gamma *= 0.8
            self._trial(gamma, rec["fused"])
            st.allreduce(st.kl2[1:2])
            kl_new, ld_new = torch.stack([st.kl2[1], st.ld2[1]]).tolist()
            of_value = kl_new + self.lam * ld_new
        self._gamma = min(1.0, 1.2 * gamma)
        st.W, st.W_trial = st.W_trial, st.W
        st.H, st.H_trial = st.H_trial, st.H
        self.launch_stats["back_tracked_iterations"] += 1
        return False

    def _fit_loop_run_ahead(self, n_given: int, verbose, verbosity_freq):
        """The reference loop (signature_nmf.py:361-380 around mvnmf.py:197-210) without a host round trip per iteration.
        The reference looks at the objective after the first line-search trial to decide whether to back-track
        (mvnmf.py:69-92); here the H step and the unconstrained W step of iteration n + 1 are already queued behind iteration
        n when the host reads n's decision -- by then it has long arrived -- and only a rejected full step (rare: gamma adapts)
        costs a synchronisation and the two wasted passes.  Same iterates, objective history and gamma as the reference."""
        st = self._dev
        ring = {"host": torch.zeros((4, 4), dtype=torch.float64).pin_memory(), "events": [torch.cuda.Event() for _ in range(4)], "next": 0}
        self.launch_stats = {"graphs": 0, "driver": "run-ahead line search", "back_tracked_iterations": 0}
        of_values = [self.objective_function()]
        n_iteration, converged, pending = 0, False, None
        h_step_done = False    # the accepted trial already applied this iteration's H step (fused pass)
        logdet_known = False   # ld2[1] holds ln det of the current W (left there by the trial that produced it)
        freq, max_it = int(self.conv_test_freq), int(self.max_iterations)
        while not converged:
            n_iteration += 1
            if verbose and n_iteration % verbosity_freq == 0:
                print(f"iteration: {n_iteration}; objective: {of_values[-1]:.2f}")
            if not h_step_done:
                self._update_H()
            self._update_W_unconstrained(n_given, logdet_known, with_first_trial=True)
            if pending is not None and not self._confirm(pending, ring):
                # iteration n - 1 had to back-track: what was just queued started from the wrong iterate
                if not pending["fused"]:
                    self._update_H()
                self._update_W_unconstrained(n_given, True, with_first_trial=True)
            # the H step of iteration n + 1 rides on this iteration's trial pass -- unless the iterate itself is needed next:
            # for the objective of the convergence test or as the final state
            fused = bool(self.fuse_h_step) and n_iteration % freq != 0 and n_iteration < max_it
            pending = self._first_trial(ring, fused)
            h_step_done, logdet_known = fused, True
            if n_iteration % freq == 0 or n_iteration >= max_it:
                self._confirm(pending, ring)  # the objective below (and the final state) need the true iterate
                pending = None
            if n_iteration % freq == 0:
                prev = of_values[-1]
                of_values.append(self.objective_function())
                rel_change = np.abs(prev - of_values[-1]) / np.abs(prev)
                converged = bool(rel_change < self.tol and n_iteration >= self.min_iterations)
            converged |= n_iteration >= max_it
        return of_values, n_iteration

    def _fit_loop_small(self, n_given: int, verbose, verbosity_freq):
        """One launch of the persistent single-CTA kernel (sal_mvnmf_small_updates) per convergence-test period: the penalised
        objective of the period's incoming iterate plus its ``conv_test_freq`` iterations, line search and gamma included.
        The next period is launched before the host looks at that objective; states (W, H, gamma) rotate through three
        buffers, so the one the convergence test may fall back to is never overwritten.  Same iterates, history and
        stopping iteration as the reference loop (signature_nmf.py:361-380)."""
        st = self._dev
        freq, max_it, min_it = int(self.conv_test_freq), int(self.max_iterations), int(self.min_iterations)
        Wb = [st.W, torch.empty_like(st.W), torch.empty_like(st.W)]
        Hb = [st.H, torch.empty_like(st.H), torch.empty_like(st.H)]
        gam = torch.full((3,), float(self._gamma), dtype=torch.float64, device=st.device)
        obj_dev = torch.zeros(2, dtype=torch.float64, device=st.device)
        obj_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        events = [torch.cuda.Event(), torch.cuda.Event()]

        def n_of(i: int) -> int:
            return min(i * freq, max_it)

        def launch(i: int) -> None:  # state i -> state i + 1, objective of state i
            n_i, s = n_of(i), i % 2
            a, b = i % 3, (i + 1) % 3
            st.ws.mvnmf_small_updates(
                st.X, Wb[a], Wb[b], Hb[a], Hb[b], self.lam, self.delta, n_given, min(freq, max_it - n_i),
                gam[a : a + 1], gam[b : b + 1], objective=obj_dev[s : s + 1],
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
        self._gamma = float(gam[final % 3].item())
        self.launch_stats = {"graphs": 0, "periods": i + 1, "driver": "single-CTA persistent kernel"}
        return of_values, n_of(final)

    def _setup_fitting_parameters(self, fitting_kwargs: dict[str, Any] | None = None) -> None:
        self._gamma = 1.0
