"""
``StandardNMF``: models parameterised by a signature matrix W and an exposure matrix H
(reference models/standard_nmf.py:18-58).  Holds the device-resident state shared by
KLNMF and MvNMF and the sample sharding across ranks.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _dist
from .._device import PASS_SAMPLEWISE, Workspace
from ..initialization.initialize import initialize_standard_nmf
from .signature_nmf import SignatureNMF


class StandardState:
    """X [D_local][V], W [k][V], H [D_local][k] in HBM plus scratch; D_local = this rank's row block."""

    def __init__(self, model: "StandardNMF"):
        dev = model._resolved_device()
        dt = model.dtype
        model.transfer_bytes = {"h2d": 0, "d2h": 0}
        self.model = model
        self.device, self.dtype = dev, dt
        X_host = np.asarray(model.adata.X)
        W_host = np.asarray(model.asignatures.X, dtype=np.float64)
        H_dev = model.adata.obsm["exposures"] if isinstance(model.adata.obsm["exposures"], torch.Tensor) else None
        H_host = None if H_dev is not None else np.asarray(model.adata.obsm["exposures"])
        rc = getattr(model, "_resident_counts", None)
        self.k = W_host.shape[0]
        self.rank, self.world = (0, 1) if model.replica else _dist.world()
        if model.shard_input:
            self.D_total, self.V = X_host.shape
            self.lo, self.hi = _dist.shard_bounds(self.D_total, self.world, self.rank)
            if self.world > 1:  # replicas must start bit-identical whatever the host RNG did
                W_host = _dist.broadcast_numpy(W_host, dev)
                if H_host is not None:
                    H_host = _dist.broadcast_numpy(H_host, dev)
        else:  # adata already holds this rank's rows only
            self.V = X_host.shape[1]
            self.lo, self.hi = 0, X_host.shape[0]
            self.D_total = self.hi
            if self.world > 1:
                W_host = _dist.broadcast_numpy(W_host, dev)
        D = self.hi - self.lo
        self.ws = Workspace(self.V, D, self.k, dt, dev, math=model.math)
        self.X = rc.X if rc is not None else self.upload(X_host[self.lo : self.hi])  # (resident: shared by the fits of a sweep)
        if model._clip_on_device:
            changed = torch.zeros(1, dtype=torch.int64, device=dev)
            self.ws.clip_counts(self.X, changed)
            if model.shard_input:
                self.allreduce(changed)  # every rank must take the same branch: the gather below is a collective
            if int(changed.item()) > 0:  # reference signature_nmf.py:281 rebinds adata.X to the clipped matrix
                model.adata.X = self.rows_to_host(self.X)
            model._clip_on_device = False
        self.W = self.upload(W_host)
        self.H = H_dev[self.lo : self.hi].to(dt).contiguous() if H_dev is not None else self.upload(H_host[self.lo : self.hi])
        scale = getattr(model, "_exposure_scale", None)
        if scale is not None:
            self.ws.scale_clip_rows(self.H, self.upload(np.asarray(scale, dtype=np.float64)))
            model._exposure_scale = None
        self.W_next = torch.empty_like(self.W)  # the joint update writes W here (H needs the old W), then they swap
        self.Wnum = torch.zeros((self.k, self.V), dtype=dt, device=dev)
        self.obj = torch.zeros(1, dtype=torch.float64, device=dev)
        self.weights: dict[str, Any] = {}
        self.fit_loop: dict[str, Any] = {}  # spare buffers / CUDA graphs of the period-wise fit driver (KLNMF._fit_loop)

    _STAGED_DOWNLOAD_MIN_BYTES = 1 << 22

    def upload(self, host) -> torch.Tensor:
        """Host array -> contiguous device tensor of the model dtype (async DMA when the array is pinned)."""
        t = torch.from_numpy(np.ascontiguousarray(host))
        self.model.transfer_bytes["h2d"] += t.numel() * t.element_size()
        return t.to(self.device, non_blocking=True).to(self.dtype).contiguous()

    def download(self, t: torch.Tensor) -> np.ndarray:
        """Device tensor -> float64 host array (the dtype the reference leaves in the AnnData objects).

        Large results: widened to float64 ON THE DEVICE and copied by DMA straight into a page-locked array that becomes the
        result (numpy view of a pinned torch tensor; torch's pinned-host cache recycles the block once the array is dropped):
        160 MB at PCIe rate instead of 80 MB + a host pass over 240 MB.  `model.pinned_results = False` restores the staged
        path (fp32 DMA into a cached pinned buffer, multi-threaded widening copy into a pageable array)."""
        out = None
        if t.numel() * 8 >= self._STAGED_DOWNLOAD_MIN_BYTES:
            try:
                if getattr(self.model, "pinned_results", True):
                    res = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
                    res.copy_(t.to(torch.float64), non_blocking=True)
                    torch.cuda.current_stream(self.device).synchronize()
                    out = res.numpy()
                    moved = res.numel() * 8
                else:
                    stage = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    stage.copy_(t.contiguous(), non_blocking=True)
                    torch.cuda.current_stream(self.device).synchronize()
                    out = torch.empty(t.shape, dtype=torch.float64).copy_(stage).numpy()
                    moved = stage.numel() * stage.element_size()
            except RuntimeError:
                out = None  # page-locked memory exhausted: plain pageable copy below
        if out is None:
            out = t.to(torch.float64).cpu().numpy()
            moved = out.nbytes
        self.model.transfer_bytes["d2h"] += moved  # bytes that crossed the bus
        return out

    def shard(self, per_sample) -> torch.Tensor | None:
        """Upload this rank's slice of a per-sample host vector (or None)."""
        if per_sample is None:
            return None
        return self.upload(np.asarray(per_sample, dtype=np.float64)[self.lo : self.hi])

    def rows_to_host(self, local: torch.Tensor) -> np.ndarray:
        """Per-sample device rows -> host array covering ``adata``'s rows (all-gather when the model sharded them)."""
        if self.model.shard_input and self.world > 1:
            local = _dist.gather_rows(local, self.D_total)
        return self.download(local)

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        """Sum over the ranks that share this fit (no-op for a single GPU or an independent replica)."""
        if self.world > 1:
            _dist.allreduce_sum_(t)
        return t

    def objective_value(self) -> float:
        """Sum the device scalar over ranks and bring it to the host (one sync)."""
        self.allreduce(self.obj)
        return float(self.obj.item())

    def close(self) -> None:
        self.ws.close()


class StandardNMF(SignatureNMF):
    def _initialize(self, given_parameters=None, init_kwargs=None) -> None:
        """Initialise signatures and exposures; given signatures are kept fixed (reference :32-58)."""
        init_kwargs = {} if init_kwargs is None else init_kwargs.copy()
        init_kwargs.update(self._init_device_kwargs())
        defer: dict[str, Any] = {}
        self.asignatures = initialize_standard_nmf(
            self.adata, self.n_signatures, self.init_method, given_parameters, _defer=defer, **init_kwargs
        )
        # large custom exposures: H <- clip(H * colsum(W)) (reference initialize.py:116-118) runs on the device
        self._exposure_scale = defer.get("exposure_scale")

    def _to_device(self) -> None:
        self._dev = StandardState(self)
        self._upload_fitting_parameters()

    def _upload_fitting_parameters(self) -> None:
        """Hook for per-sample weights etc."""

    def _to_host(self) -> None:
        st = self._dev
        keep = getattr(self, "_download_if", None)
        if keep is not None and not keep(self):  # a sweep's fit that is not the best of its k: only the error is kept
            self.adata.obsm["exposures"] = None
            return
        self.asignatures.X = st.download(st.W)
        self.adata.obsm["exposures"] = st.rows_to_host(st.H)

    def compute_reconstruction_errors(self) -> None:
        """Unweighted per-sample KL divergences -> ``adata.obs['reconstruction_error']``
        (reference klnmf.py:54-62, mvnmf.py:139-147 -> _utils_klnmf.py:58-97)."""
        with self._resident() as st:
            out = torch.empty(st.hi - st.lo, dtype=st.dtype, device=st.device)
            st.ws.klnmf_pass(st.X, st.W, st.H, PASS_SAMPLEWISE, per_sample=out)
            self.adata.obs["reconstruction_error"] = np.array(st.rows_to_host(out))

    @staticmethod
    def _n_given(given_parameters) -> int:
        if given_parameters and "asignatures" in given_parameters:
            return given_parameters["asignatures"].n_obs
        return 0
