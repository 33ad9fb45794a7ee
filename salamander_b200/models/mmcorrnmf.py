"""
``MultimodalCorrNMF``: several correlated-NMF models, one per data modality, that share the sample embeddings (and
the prior variance).  Same interface as the numerics / fit part of reference models/mmcorrnmf.py:34-491
(constructor, ``fit(mdata, given_parameters, init_kwargs, history, verbose, verbosity_freq)``, the per-modality and
joint update methods the reference's tests call); the plotting wrappers (:493-739) are out of scope.

Per modality the device kernels are those of CorrNMFDet (one pass over that modality's X per iteration for aux and
the W numerator); the joint sample-embedding problem concatenates all modalities' signature embeddings, scalings
and aux rows and uses the per-modality sample scalings as a vector scaling (sal_corrnmf_sample_embeddings_mm).
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np
import pandas as pd
import torch

from .._anndata import AnnData, MuData
from .._device import PASS_NOCLIP, PASS_POISSON, PASS_SAMPLEWISE, PASS_UPDATE_H, PASS_WNUM, Workspace, resolve_device, resolve_dtype
from ..initialization.initialize import initialize_mmcorrnmf
from ..initialization.methods import _INIT_METHODS
from ..utils import EPSILON, type_checker, value_checker
from .corrnmf import CorrState, signature_embeddings_update


class _ModView:
    """What CorrState needs from a model, for one modality."""

    def __init__(self, parent: "MultimodalCorrNMF", mod_name: str, k: int):
        self.adata, self.asignatures = parent.mdata[mod_name], parent.asignatures[mod_name]
        self.n_signatures, self.dim_embeddings = k, parent.dim_embeddings
        self.dtype, self._clip_on_device = parent.dtype, False
        self.transfer_bytes = parent.transfer_bytes
        self.replica = getattr(parent, "replica", False)
        self._parent = parent

    def _resolved_device(self):
        return resolve_device(self._parent.device)


class _MMState:
    def __init__(self, model: "MultimodalCorrNMF"):
        self.mods: dict[str, CorrState] = {}
        U_host = np.asarray(model.mdata.obsm["embeddings"], dtype=np.float64)
        for mod_name, k in zip(model.mod_names, model.ns_signatures):
            adata = model.mdata[mod_name]
            adata.obsm["embeddings"] = U_host  # CorrState uploads it; replaced by the shared tensor below
            # several GPUs: samples sharded by contiguous row blocks, identically for every modality; the
            # per-modality parameters W / a / L and the variance are replicated, b / H / aux and the SHARED sample embeddings
            # are this rank's rows.  ``model.shard = False`` keeps every rank a replica of the whole problem instead.
            st = CorrState(_ModView(model, mod_name, k), allow_shard=bool(getattr(model, "shard", True)))
            del adata.obsm["embeddings"]
            self.mods[mod_name] = st
        first = next(iter(self.mods.values()))
        self.device, self.dtype, self.D, self.m = first.device, first.dtype, first.D, first.m  # D: this rank's samples
        self.D_total, self.world = first.D_total, first.world
        model.shard_info = {"world": first.world, "rank": first.rank, "local_samples": first.D, "samples": first.D_total}
        self.U = first.U
        for st in self.mods.values():
            st.U = self.U
        self.K = int(sum(model.ns_signatures))
        if self.K > 32:
            raise NotImplementedError("more than 32 signatures over all modalities are not supported by the joint sample-embedding kernel")
        self.ws_joint = Workspace(1, self.D, self.K, self.dtype, self.device, math="fma")
        self.norms = torch.zeros(3, dtype=torch.float64, device=self.device)

    def close(self) -> None:
        for st in self.mods.values():
            st.close()
        self.ws_joint.close()


class MultimodalCorrNMF:
    def __init__(
        self,
        ns_signatures: list[int],
        dim_embeddings: int | None = None,
        init_method: str = "nndsvd",
        min_iterations: int = 500,
        max_iterations: int = 10000,
        conv_test_freq: int = 10,
        tol: float = 1e-7,
        *,
        device=None,
        dtype="float64",
    ):
        value_checker("init_method", init_method, _INIT_METHODS)
        self.ns_signatures = list(ns_signatures)
        self.dim_embeddings = int(np.max(ns_signatures)) if dim_embeddings is None else dim_embeddings
        self.init_method = init_method
        self.min_iterations, self.max_iterations = min_iterations, max_iterations
        self.conv_test_freq, self.tol = conv_test_freq, tol
        self.variance = 1.0
        self.device, self.dtype = device, resolve_dtype(dtype)
        names = [f"mod{n}" for n in range(1, len(ns_signatures) + 1)]
        self.mdata = MuData({name: AnnData() for name in names})
        self.asignatures = {name: AnnData() for name in names}
        self.history: dict[str, Any] = {}
        self.transfer_bytes = {"h2d": 0, "d2h": 0}
        self._dev: _MMState | None = None
        self._in_fit = False
        self.n_iterations = 0

    # ---- accessors (reference :72-113) -------------------------------------------------------------
    @property
    def mod_names(self) -> list[str]:
        return list(self.mdata.mod.keys())

    @property
    def signature_names(self) -> dict[str, list[str]]:
        return {name: list(asigs.obs_names) for name, asigs in self.asignatures.items()}

    @property
    def sample_names(self) -> list[str]:
        return list(self.mdata.obs_names)

    @property
    def signatures(self) -> dict[str, pd.DataFrame]:
        return {name: asigs.to_df() for name, asigs in self.asignatures.items()}

    @property
    def exposures(self) -> dict[str, pd.DataFrame]:
        return {
            name: pd.DataFrame(self.mdata[name].obsm["exposures"], index=self.sample_names, columns=self.asignatures[name].obs_names)
            for name in self.mod_names
        }

    @property
    def objective(self) -> Literal["minimize", "maximize"]:
        return "maximize"

    # ---- device residency ----------------------------------------------------------------------------
    class _Resident:
        def __init__(self, model):
            self.m, self.owner = model, False

        def __enter__(self):
            if self.m._dev is None:
                self.m.transfer_bytes = {"h2d": 0, "d2h": 0}
                self.m._dev = _MMState(self.m)
                self.owner = True
            return self.m._dev

        def __exit__(self, exc_type, exc, tb):
            if self.owner:
                try:
                    if exc_type is None:
                        self.m._to_host()
                finally:
                    self.m._dev.close()
                    self.m._dev = None
            return False

    def _resident(self):
        return MultimodalCorrNMF._Resident(self)

    def _to_host(self) -> None:
        dev = self._dev
        for name, st in dev.mods.items():
            asigs, adata = self.asignatures[name], self.mdata[name]
            asigs.X = st.download(st.W)
            asigs.obs["scalings"] = st.download(st.a)
            asigs.obsm["embeddings"] = st.download(st.L)
            adata.obs["scalings"] = st.rows_to_host(st.b)
            adata.obsm["exposures"] = st.rows_to_host(st.H)
        self.mdata.obsm["embeddings"] = next(iter(dev.mods.values())).rows_to_host(dev.U)

    # ---- numerics (reference :106-115, :168-194, :233-453) ---------------------------------------------
    def compute_exposures(self) -> None:
        with self._resident() as dev:
            for st in dev.mods.values():
                st.call("sal_corrnmf_exposures", st.a, st.b, st.L, st.U, st.m, st.H)

    def compute_reconstruction_errors(self) -> None:
        with self._resident() as dev:
            for name, st in dev.mods.items():
                st.call("sal_corrnmf_exposures", st.a, st.b, st.L, st.U, st.m, st.H)
                out = torch.empty(st.D, dtype=st.dtype, device=st.device)
                st.ws.klnmf_pass(st.X, st.W, st.H, PASS_SAMPLEWISE, per_sample=out)
                self.mdata[name].obs["reconstruction_error"] = st.rows_to_host(out)

    @property
    def reconstruction_errors(self) -> dict[str, float]:
        if any("reconstruction_error" not in adata.obs for adata in self.mdata.mod.values()):
            self.compute_reconstruction_errors()
        return {name: float(np.sum(adata.obs["reconstruction_error"])) for name, adata in self.mdata.mod.items()}

    @property
    def reconstruction_error(self) -> float:
        return float(np.sum(list(self.reconstruction_errors.values())))

    def objective_function(self) -> float:
        """ELBO: modality ELBOs without the sample-embedding prior, which is counted once (reference :168-194)."""
        with self._resident() as dev:
            var, m = float(self.variance), dev.m
            elbo, sumU2 = 0.0, 0.0
            for st in dev.mods.values():
                st.ws.klnmf_pass(st.X, st.W, st.H, PASS_POISSON, objective=st.obj)
                st.allreduce(st.obj)
                need = st.lgamma_sum is None
                norms = st.norms_all(need)  # [sum L^2, sum U^2 and sum lnGamma(1 + x) over ALL samples]
                if need:
                    st.lgamma_sum = norms[2]
                elbo += float(st.obj.item()) - st.lgamma_sum - 0.5 * m * st.k * np.log(2 * np.pi * var) - norms[0] / (2 * var)
                sumU2 = norms[1]
            elbo -= 0.5 * m * dev.D_total * np.log(2 * np.pi * var) + sumU2 / (2 * var)
            return float(elbo)

    def _compute_auxs(self):
        with self._resident() as dev:
            for st in dev.mods.values():
                st.ws.klnmf_pass(st.X, st.W, st.H, PASS_UPDATE_H | PASS_WNUM | PASS_NOCLIP, H_out=st.auxT, Wnum=st.Wnum)
                st.allreduce(st.Wnum)
            return None if self._in_fit else {name: st.rows_to_host(st.auxT).T for name, st in dev.mods.items()}

    @staticmethod
    def _aux_up(st, aux) -> None:
        if aux is not None:  # (k, n_samples) host array over ALL samples, like the reference's
            st.auxT = st.upload(np.asarray(aux, dtype=np.float64).T[st.lo : st.hi])

    def update_sample_scalings(self, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        with self._resident() as dev:
            for name, st in dev.mods.items():
                if "sample_scalings" not in given_parameters.get(name, {}):
                    st.call("sal_corrnmf_sample_scalings", st.xsum, st.a, st.L, st.U, st.m, st.b)

    def update_signature_scalings(self, auxs=None, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        with self._resident() as dev:
            for name, st in dev.mods.items():
                self._aux_up(st, None if auxs is None else auxs[name])
                if "signature_scalings" not in given_parameters.get(name, {}):
                    st.call("sal_corrnmf_signature_scalings_sums", st.auxT, st.b, st.L, st.U, st.m, st.sums)
                    st.allreduce(st.sums)
                    st.call("sal_corrnmf_signature_scalings_finish", st.sums, st.a)

    def update_variance(self, given_parameters: dict[str, Any] | None = None) -> None:
        if given_parameters and "variance" in given_parameters:
            return
        with self._resident() as dev:
            total, count = 0.0, 0
            for st in dev.mods.values():
                sumL2, sumU2 = st.norms_all(False)[:2]
                total += sumL2
                count += st.k * st.m
            total += sumU2
            count += dev.D_total * dev.m
            self.variance = float(np.clip(total / count, EPSILON, None))

    def update_signatures(self, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        with self._resident() as dev:
            for name, st in dev.mods.items():
                gm = given_parameters.get(name, {})
                n_given = gm["asignatures"].n_obs if "asignatures" in gm else 0
                if not self._in_fit:
                    st.ws.klnmf_pass(st.X, st.W, st.H, PASS_WNUM, Wnum=st.Wnum)
                    st.allreduce(st.Wnum)
                st.ws.w_epilogue(st.W, st.Wnum, n_given, False, st.W)

    def update_signature_embeddings(self, auxs=None, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        with self._resident() as dev:
            for name, st in dev.mods.items():
                self._aux_up(st, None if auxs is None else auxs[name])
                if "signature_embeddings" not in given_parameters.get(name, {}):
                    signature_embeddings_update(st, float(self.variance))

    def update_sample_embeddings(self, auxs=None) -> None:
        with self._resident() as dev:
            sts = list(dev.mods.values())
            for name, st in dev.mods.items():
                self._aux_up(st, None if auxs is None else auxs[name])
            aux_all = torch.cat([st.auxT for st in sts], dim=1).contiguous()
            a_all = torch.cat([st.a for st in sts]).contiguous()
            L_all = torch.cat([st.L for st in sts], dim=0).contiguous()
            b_mat = torch.cat([st.b[:, None].expand(st.D, st.k) for st in sts], dim=1).contiguous()
            st0 = sts[0]
            import ctypes as C

            from .. import _lib

            stream = C.c_void_p(torch.cuda.current_stream(dev.device).cuda_stream)
            p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
            _lib.check(
                st0.lib.sal_corrnmf_sample_embeddings_mm(
                    dev.ws_joint._h, p(aux_all), p(a_all), p(b_mat), p(L_all), p(dev.U), dev.m, float(self.variance), 3, stream
                ),
                "sal_corrnmf_sample_embeddings_mm",
            )

    def update_embeddings(self, auxs=None, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        self.update_signature_embeddings(auxs, given_parameters)
        if "sample_embeddings" not in given_parameters:
            self.update_sample_embeddings(auxs)

    def _update_parameters(self, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        with self._resident():
            in_fit, self._in_fit = self._in_fit, True
            try:
                self.update_sample_scalings(given_parameters)
                self.compute_exposures()
                self._compute_auxs()
                self.update_signature_scalings(None, given_parameters)
                self.update_embeddings(None, given_parameters)
                self.update_variance(given_parameters)
                self.update_signatures(given_parameters)
            finally:
                self._in_fit = in_fit

    # ---- fit (reference :196-231, :455-491) ------------------------------------------------------------
    def _setup_mdata(self, mdata) -> None:
        type_checker("mdata", mdata, MuData)
        if mdata.n_mod != len(self.ns_signatures):
            raise ValueError(f"The data has to have {len(self.ns_signatures)} many modalities.")
        expected = list(mdata.mod.values())[0].obs_names
        for adata in mdata.mod.values():
            if not all(adata.obs_names == expected):
                raise ValueError("The sample names of the different modalities are not identical.")
        self.mdata = mdata

    def _initialize(self, given_parameters=None, init_kwargs=None) -> None:
        init_kwargs = {} if init_kwargs is None else init_kwargs.copy()
        self.asignatures, self.variance = initialize_mmcorrnmf(
            self.mdata, self.ns_signatures, self.dim_embeddings, self.init_method, given_parameters, **init_kwargs
        )
        for adata in self.mdata.mod.values():
            adata.obsm.pop("exposures", None)
        self.compute_exposures()

    def fit(self, mdata, given_parameters=None, init_kwargs=None, history: bool = True, verbose: Literal[0, 1] = 0, verbosity_freq: int = 100):
        self._setup_mdata(mdata)
        self._initialize(given_parameters, init_kwargs)
        with self._resident():
            self._in_fit = True
            try:
                of_values = [self.objective_function()]
                n_iteration, converged = 0, False
                while not converged:
                    n_iteration += 1
                    if verbose and n_iteration % verbosity_freq == 0:
                        print(f"iteration: {n_iteration}; objective: {of_values[-1]:.2f}")
                    self._update_parameters(given_parameters)
                    if n_iteration % self.conv_test_freq == 0:
                        prev = of_values[-1]
                        of_values.append(self.objective_function())
                        rel_change = np.abs(prev - of_values[-1]) / np.abs(prev)
                        converged = bool(rel_change < self.tol and n_iteration >= self.min_iterations)
                    converged |= n_iteration >= self.max_iterations
                self.n_iterations = n_iteration
            finally:
                self._in_fit = False
        if history:
            self.history["objective_function"] = of_values[1:]
        self.mdata.update()
        return self
