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

    def _setup_fitting_parameters(self, fitting_kwargs: dict[str, Any] | None = None) -> None:
        self._gamma = 1.0
