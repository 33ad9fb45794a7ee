"""
``CorrNMFDet``: correlated NMF with deterministic batch updates -- exposures are parameterised as
``exp(signature scaling + sample scaling + <signature embedding, sample embedding>)`` with Gaussian priors on the
embeddings.  Same interface as reference models/corrnmf.py:25-235 (abstract ``CorrNMF``) and
models/corrnmf_det.py:18-169 (``CorrNMFDet``): constructor, ``fit``, the single-parameter update methods the
reference's tests call, ``given_parameters`` keys that freeze individual parameters, results in
``asignatures.X / .obs['scalings'] / .obsm['embeddings']``, ``adata.obs['scalings'] / .obsm['embeddings'] /
.obsm['exposures']`` and ``model.variance``.

All numerics run on the device through the C ABI (csrc/corrnmf.cu + the fused pass): per iteration ONE pass over X
produces aux and the W numerator (both use the exposures computed before the scaling / embedding updates, like the
reference), the sample embeddings are updated one thread per sample and the signature embeddings
one CTA per signature by a device restatement of SciPy's Newton-CG (the third-party algorithm behind
_utils_corrnmf.py:400-407).
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Literal

import numpy as np
import torch

from .. import _dist, _lib
from .._device import PASS_NOCLIP, PASS_POISSON, PASS_SAMPLEWISE, PASS_UPDATE_H, PASS_WNUM, Workspace
from ..initialization.initialize import initialize_corrnmf
from .signature_nmf import SignatureNMF


class CorrState:
    """Device-resident state: X [D][V], W [k][V], a [k], b [D], L [k][m], U [D][m], H / auxT [D][k].

    Several GPUs: samples are sharded by contiguous row blocks -- X, b, U, H, aux are this rank's rows,
    W, a, L and the variance are replicated.  Sample-side updates are local; the W numerator, the two k-vectors of the
    signature scalings, sum |U|^2 and the log-likelihood are summed over ranks (NCCL all-reduce of a few KB); for the
    signature embeddings the per-sample inputs (aux, b, U) are all-gathered and every rank runs the Newton-CG of ITS
    share of the signatures on all samples, so that step is divided over the GPUs by signature instead of by sample and
    needs no collective inside the solver."""

    def __init__(self, model: "CorrNMF", allow_shard: bool = True):
        dev = model._resolved_device()
        dt = model.dtype
        self.rank, self.world = (0, 1) if getattr(model, "replica", False) else _dist.world()
        if self.world > 1 and not allow_shard:  # replicas: every rank holds and fits the whole problem
            self.rank, self.world = 0, 1
        model.transfer_bytes = {"h2d": 0, "d2h": 0}
        self.model, self.device, self.dtype = model, dev, dt
        X = np.asarray(model.adata.X)
        self.D_total, self.V = X.shape
        self.lo, self.hi = _dist.shard_bounds(self.D_total, self.world, self.rank)
        self.D = self.hi - self.lo
        self.k, self.m = int(model.n_signatures), int(model.dim_embeddings)
        self.ws = Workspace(self.V, self.D, self.k, dt, dev, math="fma")
        # the gathered problem of the signature embeddings (all samples) has its own handle
        self.ws_full = Workspace(self.V, self.D_total, self.k, dt, dev, math="fma") if self.world > 1 else self.ws
        self.lib = self.ws.lib
        if self.m > self.lib.sal_corrnmf_max_dim():
            raise NotImplementedError(f"dim_embeddings > {self.lib.sal_corrnmf_max_dim()} is not supported by the device kernels.")
        up = self.upload
        rows = slice(self.lo, self.hi)
        self.X = up(X[rows])
        if model._clip_on_device:
            changed = torch.zeros(1, dtype=torch.int64, device=dev)
            self.ws.clip_counts(self.X, changed)
            self.allreduce(changed)
            if int(changed.item()) > 0:
                model.adata.X = self.rows_to_host(self.X)
            model._clip_on_device = False
        asig, adata = model.asignatures, model.adata

        def host(arr):  # replicas must start bit-identical whatever the host initialisation did
            arr = np.asarray(arr, dtype=np.float64)
            return _dist.broadcast_numpy(arr, dev) if self.world > 1 else arr

        self.W = up(host(asig.X))
        self.a = up(host(asig.obs["scalings"].values))
        self.L = up(host(asig.obsm["embeddings"]))
        self.b = up(host(adata.obs["scalings"].values)[rows])
        self.U = up(host(adata.obsm["embeddings"])[rows])
        if "exposures" in adata.obsm:
            self.H = up(host(adata.obsm["exposures"])[rows])
        else:
            self.H = torch.empty((self.D, self.k), dtype=dt, device=dev)
        self.auxT = torch.empty((self.D, self.k), dtype=dt, device=dev)
        self.Wnum = torch.empty((self.k, self.V), dtype=dt, device=dev)
        self.xsum = torch.empty(self.D, dtype=dt, device=dev)
        self.sums = torch.zeros(2 * self.k, dtype=torch.float64, device=dev)
        self.norms = torch.zeros(3, dtype=torch.float64, device=dev)
        self.obj = torch.zeros(1, dtype=torch.float64, device=dev)
        self.call("sal_row_sums", self.X, self.xsum)
        self.lgamma_sum = None  # sum lnGamma(1 + x), constant of the ELBO, computed at the first objective

    # -- ranks -----------------------------------------------------------------------------------
    def sig_exchange(self):
        """Symmetric receive buffers of the signature-embedding solver's in-kernel exchange (``None``: single GPU, or peer
        memory is unavailable on at least one rank -- decided collectively, once per device state)."""
        if self.world == 1 or getattr(self.model, "allreduce", "auto") == "nccl":
            return None
        if not hasattr(self, "_sig_px"):
            px = None
            try:
                nbytes = int(self.lib.sal_corrnmf_sig_exchange_bytes(32, self.world))  # sized for the largest k: shared by all fits
                px = _dist.shared_peer_exchange(nbytes, self.device, name="corrnmf_sig")
            except Exception:  # pragma: no cover - depends on the system
                px = None
            self._sig_px = px if _dist.all_ranks_agree(px is not None, self.device) else None
        return self._sig_px

    def small_exchange(self):
        """Receive buffers of the peer-memory all-reduce of the iteration's few-KB sums (``None``: single GPU or peer memory
        unavailable on at least one rank -- decided collectively, once per device state; NCCL is used then)."""
        if self.world == 1 or getattr(self.model, "allreduce", "auto") == "nccl":
            return None
        if not hasattr(self, "_small_px"):
            px = None
            try:
                px = _dist.shared_peer_exchange(int(self.lib.sal_p2p_allreduce_bytes(self.world)), self.device, name="small_allreduce")
            except Exception:  # pragma: no cover - depends on the system
                px = None
            self._small_px = px if _dist.all_ranks_agree(px is not None, self.device) else None
            self._small_max = int(self.lib.sal_p2p_allreduce_max_values())
        return self._small_px

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a small device tensor.  float64 sums of up to a few thousand values go through peer
        memory (sal_p2p_allreduce_f64: one small kernel, contributions added in rank order, ~5 us instead of a library
        collective's ~20); anything else through NCCL."""
        if self.world > 1:
            px = self.small_exchange()
            if px is not None and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self._small_max:
                self.call("sal_p2p_allreduce_f64", t, t.numel(), px.peers, self.world, self.rank, _dist.next_launch_id(px))
            else:
                _dist.allreduce_sum_(t)
        return t

    def rows_to_host(self, local: torch.Tensor) -> np.ndarray:
        """Per-sample device rows -> host array over all samples (all-gather when sharded)."""
        return self.download(_dist.gather_rows(local, self.D_total) if self.world > 1 else local)

    def norms_all(self, with_lgamma: bool) -> list[float]:
        """[sum L^2, sum U^2 over all samples, sum lnGamma(1 + X) over all samples]"""
        self.call("sal_corrnmf_norms", self.L, self.U, self.m, self.X if with_lgamma else None, self.norms)
        if self.world > 1:
            if self.rank != 0:
                self.norms[0:1] = 0  # L is replicated: count it once
            self.allreduce(self.norms)
        return self.norms.tolist()

    # -- plumbing ------------------------------------------------------------------------------
    def upload(self, host) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(host))
        self.model.transfer_bytes["h2d"] += t.numel() * t.element_size()
        return t.to(self.device, non_blocking=True).to(self.dtype).contiguous()

    def download(self, t: torch.Tensor) -> np.ndarray:
        out = t.to(torch.float64).cpu().numpy()
        self.model.transfer_bytes["d2h"] += out.nbytes
        return out

    def call(self, name: str, *args, ws: Workspace | None = None) -> None:
        """Call ``name(handle, *args, stream)``: tensors become device pointers, ints / floats pass through."""
        conv = []
        for x in args:
            if isinstance(x, torch.Tensor):
                conv.append(C.c_void_p(x.data_ptr()))
            elif x is None:
                conv.append(None)
            else:
                conv.append(x)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(getattr(self.lib, name)((ws or self.ws)._h, *conv, stream), name)

    def close(self) -> None:
        if self.ws_full is not self.ws:
            self.ws_full.close()
        self.ws.close()


class CorrNMF(SignatureNMF):
    def __init__(
        self,
        n_signatures: int = 1,
        init_method: str = "nndsvd",
        dim_embeddings: int | None = None,
        min_iterations: int = 500,
        max_iterations: int = 10000,
        conv_test_freq: int = 10,
        tol: float = 1e-7,
        **device_kwargs,
    ):
        device_kwargs.pop("math", None)  # the correlated models always use the exact kernels
        super().__init__(n_signatures, init_method, min_iterations, max_iterations, conv_test_freq, tol, **device_kwargs)
        if dim_embeddings is None:
            dim_embeddings = n_signatures
        self.dim_embeddings = dim_embeddings
        self.variance = 1.0

    @property
    def objective(self) -> Literal["minimize", "maximize"]:
        return "maximize"

    # ---- device residency ------------------------------------------------------------------------
    def _to_device(self) -> None:
        self._dev = CorrState(self)

    def _to_host(self) -> None:
        st = self._dev
        self.asignatures.X = st.download(st.W)
        self.asignatures.obs["scalings"] = st.download(st.a)
        self.asignatures.obsm["embeddings"] = st.download(st.L)
        self.adata.obs["scalings"] = st.rows_to_host(st.b)
        self.adata.obsm["embeddings"] = st.rows_to_host(st.U)
        self.adata.obsm["exposures"] = st.rows_to_host(st.H)

    # ---- reference corrnmf.py:66-98 --------------------------------------------------------------
    def compute_exposures(self) -> None:
        with self._resident() as st:
            st.call("sal_corrnmf_exposures", st.a, st.b, st.L, st.U, st.m, st.H)

    def compute_reconstruction_errors(self) -> None:
        with self._resident() as st:
            st.call("sal_corrnmf_exposures", st.a, st.b, st.L, st.U, st.m, st.H)
            out = torch.empty(st.D, dtype=st.dtype, device=st.device)
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_SAMPLEWISE, per_sample=out)
            self.adata.obs["reconstruction_error"] = st.rows_to_host(out)

    def objective_function(self, penalize_sample_embeddings: bool = True) -> float:
        """The evidence lower bound with the exposures as they are stored (reference corrnmf.py:86-98 ->
        _utils_corrnmf.py:55-100): Poisson log-likelihood minus the Gaussian priors of the embeddings."""
        with self._resident() as st:
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_POISSON, objective=st.obj)
            st.allreduce(st.obj)
            need_lgamma = st.lgamma_sum is None
            norms = st.norms_all(need_lgamma)
            if need_lgamma:
                st.lgamma_sum = norms[2]
            llh, sumL2, sumU2 = float(st.obj.item()) - st.lgamma_sum, norms[0], norms[1]
            var, m = float(self.variance), st.m
            elbo = llh - 0.5 * m * st.k * np.log(2 * np.pi * var) - sumL2 / (2 * var)
            if penalize_sample_embeddings:
                elbo -= 0.5 * m * st.D_total * np.log(2 * np.pi * var) + sumU2 / (2 * var)
            return float(elbo)

    # ---- initialisation (reference corrnmf.py:104-136) -------------------------------------------
    def _initialize(self, given_parameters=None, init_kwargs=None) -> None:
        init_kwargs = {} if init_kwargs is None else init_kwargs.copy()
        init_kwargs.update(self._init_device_kwargs())
        self.asignatures, self.variance = initialize_corrnmf(
            self.adata, self.n_signatures, self.dim_embeddings, self.init_method, given_parameters, **init_kwargs
        )
        self.adata.obsm.pop("exposures", None)
        self.compute_exposures()

    def _setup_fitting_parameters(self, fitting_kwargs: dict[str, Any] | None = None) -> None:
        return


class CorrNMFDet(CorrNMF):
    # ---- single-parameter updates (reference corrnmf_det.py:27-155) ------------------------------
    def _compute_aux(self):
        """aux (and the raw W numerator) in one pass over X; returned as the (k, D) host array the reference uses."""
        with self._resident() as st:
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_UPDATE_H | PASS_WNUM | PASS_NOCLIP, H_out=st.auxT, Wnum=st.Wnum)
            st.allreduce(st.Wnum)
            return None if self._in_fit else st.rows_to_host(st.auxT).T

    def _aux_to_device(self, st, aux) -> None:
        if aux is not None:  # (k, n_samples) host array over ALL samples, like the reference's
            st.auxT = st.upload(np.asarray(aux, dtype=np.float64).T[st.lo : st.hi])

    def update_sample_scalings(self, given_parameters: dict[str, Any] | None = None) -> None:
        if given_parameters and "sample_scalings" in given_parameters:
            return
        with self._resident() as st:
            st.call("sal_corrnmf_sample_scalings", st.xsum, st.a, st.L, st.U, st.m, st.b)

    def update_signature_scalings(self, aux=None, given_parameters: dict[str, Any] | None = None) -> None:
        if given_parameters and "signature_scalings" in given_parameters:
            return
        with self._resident() as st:
            self._aux_to_device(st, aux)
            st.call("sal_corrnmf_signature_scalings_sums", st.auxT, st.b, st.L, st.U, st.m, st.sums)
            st.allreduce(st.sums)
            st.call("sal_corrnmf_signature_scalings_finish", st.sums, st.a)

    def update_variance(self, given_parameters: dict[str, Any] | None = None) -> None:
        if given_parameters and "variance" in given_parameters:
            return
        with self._resident() as st:
            sumL2, sumU2 = st.norms_all(False)[:2]
            self.variance = float(np.clip((sumL2 + sumU2) / ((st.k + st.D_total) * st.m), _lib_eps(), None))

    def update_signatures(self, given_parameters: dict[str, Any] | None = None) -> None:
        """update_W with the exposures as stored; only the non-given signatures are clipped (reference :71-86)."""
        n_given = given_parameters["asignatures"].n_obs if given_parameters and "asignatures" in given_parameters else 0
        with self._resident() as st:
            if not self._in_fit:  # standalone call: the numerator has to be computed first
                st.ws.klnmf_pass(st.X, st.W, st.H, PASS_WNUM, Wnum=st.Wnum)
                st.allreduce(st.Wnum)
            st.ws.w_epilogue(st.W, st.Wnum, n_given, False, st.W)

    def update_signature_embeddings(self, aux=None) -> None:
        with self._resident() as st:
            self._aux_to_device(st, aux)
            signature_embeddings_update(st, float(self.variance))

    def update_sample_embeddings(self, aux=None) -> None:
        with self._resident() as st:
            self._aux_to_device(st, aux)
            st.call("sal_corrnmf_sample_embeddings", st.auxT, st.a, st.b, st.L, st.U, st.m, float(self.variance), 3)

    def update_embeddings(self, aux=None, given_parameters: dict[str, Any] | None = None) -> None:
        given_parameters = given_parameters or {}
        if "signature_embeddings" not in given_parameters:
            self.update_signature_embeddings(aux)
        if "sample_embeddings" not in given_parameters:
            self.update_sample_embeddings(aux)

    def _update_parameters(self, given_parameters: dict[str, Any] | None = None) -> None:
        """One iteration in the reference's order (corrnmf_det.py:157-169)."""
        given_parameters = given_parameters or {}
        with self._resident():
            in_fit, self._in_fit = self._in_fit, True  # aux / numerator stay on the device between the steps
            try:
                self.update_sample_scalings(given_parameters)
                self.compute_exposures()
                self._compute_aux()
                self.update_signature_scalings(None, given_parameters)
                self.update_embeddings(None, given_parameters)
                self.update_variance(given_parameters)
                self.update_signatures(given_parameters)
            finally:
                self._in_fit = in_fit


def signature_embeddings_update(st: CorrState, variance: float) -> None:
    """Newton-CG update of every signature embedding of one (modality's) device state (reference corrnmf_det.py:88-113 /
    mmcorrnmf.py:337-396), whatever the number of ranks."""
    if st.world == 1:
        st.call("sal_corrnmf_signature_embeddings", st.auxT, st.a, st.b, st.L, st.U, st.m, variance)
        return
    px = st.sig_exchange()
    if px is not None:
        # every rank: Newton-CG of ALL signatures on ITS samples; the totals of each evaluation are exchanged inside the
        # kernel over NVLink and summed in rank order, so the ranks' solvers run in lock step and L stays bit-identical
        st.call(
            "sal_corrnmf_signature_embeddings_p2p", st.auxT, st.a, st.b, st.L, st.U, st.m, variance,
            px.peers, st.world, st.rank, _dist.next_launch_id(px),
        )
        return
    # fallback without peer memory -- every rank: all samples' aux / scalings / embeddings, Newton-CG for its share of the
    # signatures; the rows of L are then exchanged by summing arrays that are zero outside the owner's rows (exact)
    aux_all, b_all, U_all = (_dist.gather_rows(t, st.D_total) for t in (st.auxT, st.b, st.U))
    j0, j1 = _dist.shard_bounds(st.k, st.world, st.rank)
    st.call(
        "sal_corrnmf_signature_embeddings_range", aux_all, st.a, b_all, st.L, U_all, st.m, variance,
        j0, j1 - j0, ws=st.ws_full,
    )
    _dist.exchange_owned_rows(st.L, j0, j1)


def _lib_eps() -> float:
    return float(np.finfo(np.float32).eps)
