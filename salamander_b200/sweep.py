"""
Model-selection sweep: KLNMF fits over a range of signature counts and random restarts (BASELINE config 4; in the
reference this is a user loop, ``for k in ns: KLNMF(k).fit(adata.copy())`` followed by
``model.reconstruction_error``, tutorial.ipynb:1975-2013, restarts via ``init_method='random'`` +
``init_kwargs={'seed': s}``, tutorial.ipynb:741-742).

The (k, seed) jobs are independent: with torch.distributed initialised they are dealt round-robin to the ranks
(replicas only -- every rank holds the whole count matrix, no data-path collective) and only the
small result table is gathered at the end.  Each job is an ordinary ``KLNMF.fit`` on this rank's GPU.
"""

from __future__ import annotations

from typing import Any, Iterable

import numpy as np
import pandas as pd
import torch.distributed as dist

from . import _dist
from .models.klnmf import KLNMF


def _shallow_copy(adata):
    """A fresh container around the SAME count matrix: a fit only ever rebinds ``adata.X`` (when clipping changes an
    entry), it does not write into the array, so the fits of a sweep can share it instead of copying it every time.
    The names are shared as well: building 100,000 default string names per fit costs more than the fit's 200 updates."""
    import pandas as pd

    from ._anndata import HAVE_ANNDATA, AnnData

    if HAVE_ANNDATA:  # pragma: no cover - the real container validates its own fields
        new = AnnData(np.asarray(adata.X))
        new.obs_names, new.var_names = adata.obs_names, adata.var_names
        return new
    new = AnnData(None)
    new.X = np.asarray(adata.X)
    new._obs_names, new.var_names = adata.obs_names, adata.var_names
    new.obs = pd.DataFrame(index=adata.obs_names)
    return new


class ResidentCounts:
    """The count matrix of a sweep, uploaded and clipped ONCE and shared by all of its fits on this rank (every fit used to
    upload and clip its own copy: 38 MB and a host pass per fit at 96 x 100k, more than the fit's 200 updates take), plus the
    per-sample totals the random initialisation needs.  ``adata.X`` is rebound to the clipped host matrix once, if clipping
    changed anything (reference signature_nmf.py:281 does that in every fit)."""

    def __init__(self, adata, device, dtype):
        import torch

        from ._device import Workspace, resolve_device, resolve_dtype
        from .models.signature_nmf import EPSILON

        self.device, self.dtype = resolve_device(device), resolve_dtype(dtype)
        X = np.asarray(adata.X)
        self.shape = X.shape
        self.X = torch.from_numpy(np.ascontiguousarray(X)).to(self.device).to(self.dtype).contiguous()
        ws = Workspace(X.shape[1], X.shape[0], 1, self.dtype, self.device)
        changed = torch.zeros(1, dtype=torch.int64, device=self.device)
        ws.clip_counts(self.X, changed)
        if int(changed.item()) > 0:
            adata.X = X.clip(EPSILON)
        ws.close()
        self.host = np.asarray(adata.X)
        self.totals = self.X.sum(dim=1, dtype=torch.float64)

    def matches(self, adata, device, dtype) -> bool:
        return np.asarray(adata.X) is self.host and device == self.device and dtype == self.dtype


def sweep_klnmf(
    adata,
    ns_signatures: Iterable[int],
    n_restarts: int = 1,
    seed0: int = 0,
    keep_best: bool = True,
    **model_kwargs: Any,
) -> tuple[pd.DataFrame, dict[int, KLNMF]]:
    """Fit ``KLNMF(k, init_method='random')`` for every k in ``ns_signatures`` and every seed in
    ``seed0 .. seed0 + n_restarts - 1``.

    Returns ``(table, best)``: ``table`` has one row per fit (k, seed, n_iterations, objective,
    reconstruction_error, rank) and is identical on every rank; ``best[k]`` is the fitted model with the lowest
    reconstruction error among the restarts this rank ran (all of them in a single process).
    """
    rank, world = _dist.world()
    jobs = [(int(k), int(seed0 + r)) for k in ns_signatures for r in range(n_restarts)]
    rows, best = [], {}
    resident = None
    for idx, (k, seed) in enumerate(jobs):
        if idx % world != rank:
            continue
        model = KLNMF(n_signatures=k, init_method="random", replica=True, **model_kwargs)
        model.errors_in_fit = True
        if resident is None:  # X goes to the device once per sweep and rank
            resident = ResidentCounts(adata, model._resolved_device(), model.dtype)
        model._resident_counts = resident
        model.pinned_results = False  # kept results pile up over a sweep: pageable arrays through the cached staging buffer
        # results come to the host only for a fit that is the best of its k so far (the table needs the error alone)
        model._download_if = (lambda m, k=k: keep_best and (k not in best or m.reconstruction_error < best[k].reconstruction_error))
        model.fit(_shallow_copy(adata), init_kwargs={"seed": seed})
        err = model.reconstruction_error
        rows.append((k, seed, model.n_iterations, float(model.history["objective_function"][-1]), float(err), rank))
        if keep_best and (k not in best or err < best[k].reconstruction_error):
            best[k] = model
    if world > 1:
        gathered: list[Any] = [None] * world
        dist.all_gather_object(gathered, rows)
        rows = [r for part in gathered for r in part]
    table = pd.DataFrame(rows, columns=["n_signatures", "seed", "n_iterations", "objective", "reconstruction_error", "rank"])
    return table.sort_values(["n_signatures", "seed"]).reset_index(drop=True), best


def error_curve(table: pd.DataFrame) -> pd.Series:
    """Lowest reconstruction error per number of signatures (the curve the tutorial plots against k)."""
    return table.groupby("n_signatures")["reconstruction_error"].min()
