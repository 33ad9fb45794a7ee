"""
``SignatureNMF``: constructor, ``fit`` loop and the abstract seam of all models.

Same public surface as reference models/signature_nmf.py:138-185, 237-385 -- constructor
kwargs, ``fit(adata, given_parameters, init_kwargs, fitting_kwargs, history, verbose,
verbosity_freq)``, results inside the AnnData objects, ``history['objective_function']`` --
with two keyword-only additions: ``device`` and ``dtype`` ('float64' parity mode, 'float32'
fast mode).  Between ``_to_device`` and ``_to_host`` all model state lives in HBM and every
numerical step is a CUDA kernel of libsalamander_b200 (there is no CPU path).
Plot wrappers and dimensionality-reduction helpers are out of scope.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Literal

import numpy as np
import pandas as pd

from .._anndata import AnnData
from .._device import resolve_device, resolve_dtype
from ..initialization.methods import _INIT_METHODS
from ..utils import EPSILON, match_signatures_pair, type_checker, value_checker


class SignatureNMF(ABC):
    _DEVICE_CLIP_MIN_SIZE = 1 << 22

    def __init__(
        self,
        n_signatures: int = 1,
        init_method: str = "nndsvd",
        min_iterations: int = 500,
        max_iterations: int = 10000,
        conv_test_freq: int = 10,
        tol: float = 1e-7,
        *,
        device=None,
        dtype="float64",
        math: str = "fma",
        shard_input: bool = True,
        replica: bool = False,
        init_device: bool | str = False,
    ):
        value_checker("init_method", init_method, _INIT_METHODS)
        value_checker("init_device", init_device, (True, False, "auto"))
        value_checker("math", math, ("fma", "tf32", "tf32_always"))
        self.n_signatures = n_signatures
        self.init_method = init_method
        self.min_iterations = min_iterations
        self.max_iterations = max_iterations
        self.conv_test_freq = conv_test_freq
        self.tol = tol
        self.device = device
        self.dtype = resolve_dtype(dtype)
        self.math = math
        # multi-GPU (torch.distributed initialised): True = ``adata`` holds the whole matrix on every rank and
        # each rank takes its contiguous row block; False = ``adata`` already holds only this rank's rows.
        self.shard_input = bool(shard_input)
        # True: ignore an initialised process group -- this model is an independent replica (restarts / k-sweep)
        self.replica = bool(replica)
        # NNDSVD initialisations with the SVD on the device (initialization/device_nndsvd.py): True, False (scikit-learn's
        # randomized SVD on the host, as the reference) or "auto" (device for matrices of at least 2^22 entries).  True also
        # draws the exposures of init_method="random" on the device (same distribution, not numpy's stream)
        self.init_device = init_device
        # True: fit() also leaves the per-sample reconstruction errors in adata.obs (the reference computes them lazily on
        # first access, which here would mean a second upload of X); the k-sweep driver sets it
        self.errors_in_fit = False
        self.transfer_bytes = {"h2d": 0, "d2h": 0}  # host<->device bytes of the last fit / update

        # data / fitting dependent attributes (reference signature_nmf.py:182-185)
        self.adata = AnnData()
        self.asignatures = AnnData()
        self.history: dict[str, Any] = {}

        self._dev = None  # device-resident state while fitting
        self._clip_on_device = False
        self._exposure_scale = None
        self._in_fit = False
        self.n_iterations = 0

    # ---- accessors (reference signature_nmf.py:187-235) ----------------------------------
    @property
    def mutation_types(self) -> list[str]:
        return list(self.adata.var_names)

    @property
    def signature_names(self) -> list[str]:
        return list(self.asignatures.obs_names)

    @property
    def sample_names(self) -> list[str]:
        return list(self.adata.obs_names)

    @property
    def signatures(self) -> pd.DataFrame:
        """Signatures as an (n_features, n_signatures) frame, i.e. W."""
        return pd.DataFrame(np.asarray(self.asignatures.X).T, index=self.mutation_types, columns=self.signature_names)

    @property
    def exposures(self) -> pd.DataFrame:
        """Exposures as an (n_samples, n_signatures) frame, i.e. H^T."""
        assert "exposures" in self.adata.obsm, "Accessing the exposures requires fitting the NMF model."
        return pd.DataFrame(self.adata.obsm["exposures"], index=self.sample_names, columns=self.signature_names)

    def compute_reconstruction(self) -> None:
        """``adata.obsm['X_reconstructed'] = exposures @ signatures`` (reference signature_nmf.py:221-224); a host
        product of the downloaded factors, outside the fitting path."""
        self.adata.obsm["X_reconstructed"] = self.adata.obsm["exposures"] @ self.asignatures.X

    @property
    def data_reconstructed(self) -> pd.DataFrame:
        if "X_reconstructed" not in self.adata.obsm:
            self.compute_reconstruction()
        return pd.DataFrame(self.adata.obsm["X_reconstructed"], index=self.sample_names, columns=self.mutation_types)

    def reorder(self, asignatures_other: AnnData, metric: str = "cosine", keep_names: bool = False) -> None:
        """Reorder signatures and exposure columns to match another collection of signatures
        (reference signature_nmf.py:387-406) -- how restarts and the models of a sweep are aligned."""
        names = self.asignatures.obs_names
        order = match_signatures_pair(asignatures_other.to_df(), self.asignatures.to_df(), metric=metric)
        self.asignatures = self.asignatures[order, :].copy()
        self.adata.obsm["exposures"] = self.adata.obsm["exposures"][:, order]
        if not keep_names:
            self.asignatures.obs_names = names

    @abstractmethod
    def compute_reconstruction_errors(self) -> None:
        """Write per-sample reconstruction errors to ``adata.obs['reconstruction_error']``."""

    @property
    def reconstruction_error(self) -> float:
        if "reconstruction_error" not in self.adata.obs:
            self.compute_reconstruction_errors()
        return float(np.asarray(self.adata.obs["reconstruction_error"]).sum())  # (not Series.sum: 5x slower on 100k samples)

    @property
    @abstractmethod
    def objective(self) -> Literal["minimize", "maximize"]:
        ...

    @abstractmethod
    def objective_function(self) -> float:
        ...

    # ---- the seam (reference signature_nmf.py:269-313) -------------------------------------
    def _setup_adata(self, adata: AnnData) -> None:
        """Type check, then clip the counts to EPSILON *on the caller's object* (reference :269-281)."""
        type_checker("adata", adata, AnnData)
        self.adata = adata
        rc = getattr(self, "_resident_counts", None)
        if rc is not None and rc.matches(adata, self._resolved_device(), self.dtype):
            self._clip_on_device = False  # a sweep's shared, already clipped device copy (sweep.ResidentCounts)
            return
        self._resident_counts = None
        X = np.asarray(self.adata.X)
        # float64 is the parity mode: there the counts are clipped on the host BEFORE the initialisation looks at them, exactly
        # as the reference does (signature_nmf.py:280-281); in float32 mode large matrices are clipped on the device
        if str(self.dtype).endswith("float32") and X.dtype in (np.float32, np.float64) and X.size >= self._DEVICE_CLIP_MIN_SIZE:
            # large floating matrices are clipped on the device right after the upload (sal_clip_counts) and
            # the host copy is only rewritten if an entry actually changed -- same observable result
            self._clip_on_device = True
        else:
            self._clip_on_device = False
            self.adata.X = X.clip(EPSILON)

    @abstractmethod
    def _initialize(self, given_parameters=None, init_kwargs=None) -> None:
        ...

    @abstractmethod
    def _setup_fitting_parameters(self, fitting_kwargs=None) -> None:
        ...

    @abstractmethod
    def _update_parameters(self, given_parameters=None) -> None:
        ...

    # ---- device residency ------------------------------------------------------------------
    @abstractmethod
    def _to_device(self) -> None:
        """Upload X (this rank's rows), parameters and fitting weights; create the workspace."""

    @abstractmethod
    def _to_host(self) -> None:
        """Write the parameters back into the AnnData objects as float64 host arrays."""

    def _release_device(self) -> None:
        if self._dev is not None:
            tick = getattr(self, "_tick", None) or (lambda name: None)
            if getattr(self._dev, "fit_loop", None):
                self._dev.fit_loop.clear()  # captured CUDA graphs (and the NCCL work they hold) go first
            tick("release_graphs")
            self._dev.close()
            tick("release_workspace")
            self._dev = None

    def _resolved_device(self):
        return resolve_device(self.device)

    def _init_device_kwargs(self) -> dict[str, Any]:
        """``{'_init_device': torch.device}`` when this fit's NNDSVD initialisation is to run on the device."""
        if self.init_method not in ("nndsvd", "nndsvda", "nndsvdar", "random") or self.init_device is False:
            return {}
        if self.init_method == "random" and self.init_device == "auto":
            return {}  # device-drawn exposures are not numpy's draws: only on explicit request
        if self.init_device == "auto" and np.asarray(self.adata.X).size < self._DEVICE_CLIP_MIN_SIZE:
            return {}
        out = {"_init_device": self._resolved_device()}
        rc = getattr(self, "_resident_counts", None)
        if rc is not None and self.init_method == "random":
            out["_row_totals"] = rc.totals  # the exposures are drawn from the resident totals and never leave the device
        return out

    class _Resident:
        """``with self._resident():`` -- run a block with state in HBM; outside ``fit`` this uploads before
        and downloads after, which is what lets single updates be called the way the reference's tests do."""

        def __init__(self, model):
            self.m = model
            self.owner = False

        def __enter__(self):
            if self.m._dev is None:
                self.m._to_device()
                self.owner = True
            return self.m._dev

        def __exit__(self, exc_type, exc, tb):
            if self.owner:
                tick = getattr(self.m, "_tick", None) or (lambda name: None)
                try:
                    if exc_type is None:
                        self.m._to_host()
                        tick("download")
                finally:
                    self.m._release_device()
                    tick("release")
            return False

    def _resident(self):
        return SignatureNMF._Resident(self)

    def _phase_clock(self):
        """``model.profile_phases = True`` makes ``fit`` record wall-clock seconds per phase in ``model.phase_seconds``
        (with a device synchronisation at every boundary, so leave it off when timing a whole fit)."""
        self._tick = None
        if not getattr(self, "profile_phases", False):
            return lambda name: None
        import time

        import torch

        self.phase_seconds = {}
        last = [time.perf_counter()]

        def tick(name):
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            now = time.perf_counter()
            self.phase_seconds[name] = self.phase_seconds.get(name, 0.0) + now - last[0]
            last[0] = now

        self._tick = tick
        return tick  # (fit() drops self._tick again: the closure refers to the model)

    def _fit_loop(self, given_parameters, verbose, verbosity_freq) -> tuple[list[float], int]:
        """The reference's iteration / convergence logic (signature_nmf.py:361-380); returns (of_values, n)."""
        of_values = [self.objective_function()]
        n_iteration = 0
        converged = False
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
        return of_values, n_iteration

    # ---- fit (reference signature_nmf.py:315-385) ------------------------------------------
    def fit(
        self,
        adata: AnnData,
        given_parameters: dict[str, Any] | None = None,
        init_kwargs: dict[str, Any] | None = None,
        fitting_kwargs: dict[str, Any] | None = None,
        history: bool = True,
        verbose: Literal[0, 1] = 0,
        verbosity_freq: int = 1000,
    ) -> "SignatureNMF":
        tick = self._phase_clock()
        self._setup_adata(adata)
        tick("setup_adata")
        self._initialize(given_parameters, init_kwargs)
        self._setup_fitting_parameters(fitting_kwargs)
        tick("initialize")

        with self._resident():
            tick("upload")
            self._in_fit = True
            try:
                of_values, n_iteration = self._fit_loop(given_parameters, verbose, verbosity_freq)
                self.n_iterations = n_iteration
            finally:
                self._in_fit = False
            tick("fit_loop")
            if self.errors_in_fit:  # while X and the factors are still resident (a later call would upload them again)
                self.compute_reconstruction_errors()
                tick("reconstruction_errors")
        self._tick = None

        if history:
            self.history["objective_function"] = of_values[1:]
        return self
