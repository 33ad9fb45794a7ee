"""
``MvNMF``: minimum-volume NMF, KL divergence + lam * ln det(W^T W + delta I), with the
reference's H step, closed-form unconstrained W step and back-tracking line search
(reference models/mvnmf.py:95-218).  Per iteration the device does:

    pass 1  UPDATE_H                          (update_H,               _utils_klnmf.py:220-264)
    pass 2  WNUM | HSUM | OBJECTIVE on new H  (N = (X/WH) H^T, rowsums, previous objective)
    1 CTA   logdet(W), W_unconstrained        (mvnmf.py:19-24, 37-66)
    per line-search trial: 1 CTA blend/normalise/clip/logdet + pass OBJECTIVE|UPDATE_H with
            h_scale, writing the candidate H into a spare buffer that is swapped in on accept.

The reference needs >= 4 passes over X per iteration plus one per back-track; this needs 3 + 1.
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np
import torch

from .. import _dist
from .._device import PASS_HSUM, PASS_OBJECTIVE, PASS_UPDATE_H, PASS_WNUM
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

    @property
    def objective(self) -> Literal["minimize", "maximize"]:
        return "minimize"

    def _upload_fitting_parameters(self) -> None:
        st = self._dev
        st.weights["kl"] = None
        st.weights["lhalf"] = None
        st.W_unc = torch.empty_like(st.W)
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

    def _update_W_unconstrained(self, n_given_signatures: int = 0) -> None:
        """Leaves W_unconstrained in ``st.W_unc`` and the previous objective parts in kl2[0], ld2[0]."""
        st = self._dev
        st.ws.klnmf_pass(
            st.X, st.W, st.H, PASS_WNUM | PASS_HSUM | PASS_OBJECTIVE, Wnum=st.Wnum, hsum=st.hsum, objective=st.kl2[0:1]
        )
        st.allreduce(st.Wnum)
        st.allreduce(st.hsum)
        st.ws.mvnmf_logdet(st.W, self.delta, st.ld2[0:1])
        st.ws.mvnmf_w_unconstrained(st.W, st.Wnum, st.hsum, self.lam, self.delta, n_given_signatures, st.W_unc)

    def _trial(self, gamma_blend: float) -> float:
        st = self._dev
        st.ws.mvnmf_trial(st.W, st.W_unc, gamma_blend, self.delta, st.W_trial, st.h_scale, st.ld2[1:2])
        st.ws.klnmf_pass(
            st.X,
            st.W_trial,
            st.H,
            PASS_OBJECTIVE | PASS_UPDATE_H,
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

    # ---- device-side fit driver for problems that fit one SM (BASELINE config 1) ---------------------
    def _fit_loop(self, given_parameters, verbose, verbosity_freq):
        st = self._dev
        if self.use_small_kernel and st.world == 1 and not st.ws.timing and st.ws.mvnmf_small_supported():
            return self._fit_loop_small(self._n_given(given_parameters), verbose, verbosity_freq)
        return super()._fit_loop(given_parameters, verbose, verbosity_freq)

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
