"""
``KLNMF``: NMF under the (weighted) generalised KL divergence with normalised signatures and an
optional l-half sparsity penalty on the exposures.  Same interface as reference
models/klnmf.py:18-153; the numerics are the fused CUDA pass (sal_klnmf_pass) plus the
W epilogue (sal_w_epilogue).
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np

from .. import _dist
from .._device import PASS_OBJECTIVE, PASS_UPDATE_H, PASS_WNUM
from ..utils import shape_checker, type_checker
from .standard_nmf import StandardNMF

_FITTING_KWARGS = ["weights_kl", "weights_lhalf"]
_DEFAULT_FITTING_KWARGS = {kwarg: None for kwarg in _FITTING_KWARGS}


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
                _dist.allreduce_sum_(st.Wnum)
                st.ws.w_epilogue(st.W, st.Wnum, n_given, True, st.W)

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
