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
    entry), it does not write into the array, so the fits of a sweep can share it instead of copying it every time."""
    from ._anndata import AnnData

    new = AnnData(np.asarray(adata.X))
    new.obs_names, new.var_names = adata.obs_names, adata.var_names
    return new


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
    for idx, (k, seed) in enumerate(jobs):
        if idx % world != rank:
            continue
        model = KLNMF(n_signatures=k, init_method="random", replica=True, **model_kwargs)
        model.errors_in_fit = True
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
