"""
AnnData container used at the API boundary.

The reference takes and returns ``anndata.AnnData`` (models/signature_nmf.py:269-281).
anndata is an optional dependency here: when it is importable it is used unchanged,
otherwise this module provides a small stand-in with the fields the fitting path
touches: ``X, obs, obsm, obsp, obs_names, var_names, n_obs, n_vars, shape, to_df(),
copy(), __getitem__`` and a ``concat``.  Containers are out of scope of the B200 path
(SURVEY.md 2.1 L0); this is only what the boundary needs.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

try:  # pragma: no cover - not installed in the build image
    from anndata import AnnData, concat  # type: ignore

    HAVE_ANNDATA = True
except Exception:  # pragma: no cover
    HAVE_ANNDATA = False

    class AnnData:  # type: ignore[no-redef]
        def __init__(self, X=None, obs=None):
            if isinstance(X, pd.DataFrame):
                obs_names = pd.Index(X.index.astype(str))
                var_names = pd.Index(X.columns.astype(str))
                X = X.values
            elif X is not None:
                X = np.asarray(X)
                if X.ndim != 2:
                    raise ValueError("X needs to be 2-dimensional.")
                obs_names = pd.Index([str(i) for i in range(X.shape[0])])
                var_names = pd.Index([str(i) for i in range(X.shape[1])])
            else:
                obs_names = pd.Index([])
                var_names = pd.Index([])
            self.X = X
            self._obs_names = obs_names
            self.var_names = var_names
            self.obs = pd.DataFrame(index=obs_names) if obs is None else obs
            self.obsm = {}
            self.obsp = {}

        @property
        def obs_names(self):
            return self._obs_names

        @obs_names.setter
        def obs_names(self, names):
            names = pd.Index(np.asarray(names).astype(str))
            if self.X is not None and len(names) != self.n_obs:
                raise ValueError("Length of obs_names does not match n_obs.")
            self._obs_names = names
            self.obs.index = names

        @property
        def n_obs(self) -> int:
            return 0 if self.X is None else self.X.shape[0]

        @property
        def n_vars(self) -> int:
            return 0 if self.X is None else self.X.shape[1]

        @property
        def shape(self):
            return (self.n_obs, self.n_vars)

        def to_df(self) -> pd.DataFrame:
            return pd.DataFrame(self.X, index=self.obs_names, columns=self.var_names)

        def copy(self) -> "AnnData":
            new = AnnData(None if self.X is None else np.array(self.X))
            new._obs_names = self._obs_names.copy()
            new.var_names = self.var_names.copy()
            new.obs = self.obs.copy()
            new.obsm = {k: np.array(v) for k, v in self.obsm.items()}
            new.obsp = {k: np.array(v) for k, v in self.obsp.items()}
            return new

        def __getitem__(self, idx) -> "AnnData":
            rows, cols = idx if isinstance(idx, tuple) else (idx, slice(None))
            if isinstance(rows, (int, np.integer)):
                rows = slice(rows, rows + 1)
            if isinstance(cols, (int, np.integer)):
                cols = slice(cols, cols + 1)
            new = AnnData(np.asarray(self.X)[rows][:, cols])
            new._obs_names = self._obs_names[rows]
            new.var_names = self.var_names[cols]
            new.obs = self.obs.iloc[rows].copy()
            new.obsm = {k: np.asarray(v)[rows] for k, v in self.obsm.items()}
            return new

        def __repr__(self) -> str:
            return f"AnnData object with n_obs x n_vars = {self.n_obs} x {self.n_vars}"

    def concat(adatas, join="outer"):  # type: ignore[no-redef]
        adatas = list(adatas)
        out = AnnData(np.concatenate([np.asarray(a.X) for a in adatas], axis=0))
        out.var_names = adatas[0].var_names
        out.obs = pd.concat([a.obs for a in adatas], axis=0)
        out._obs_names = pd.Index(np.concatenate([np.asarray(a.obs_names) for a in adatas]))
        out.obs.index = out._obs_names
        for key in adatas[0].obsm:
            if all(key in a.obsm for a in adatas):
                out.obsm[key] = np.concatenate([np.asarray(a.obsm[key]) for a in adatas], axis=0)
        return out


try:  # pragma: no cover - not installed in the build image
    from mudata import MuData  # type: ignore

    HAVE_MUDATA = True
except Exception:  # pragma: no cover
    HAVE_MUDATA = False

    class MuData:  # type: ignore[no-redef]
        """Minimal stand-in for ``mudata.MuData``: named AnnData modalities over the same samples plus the shared
        ``obsm`` (multimodal correlated NMF keeps the joint sample embeddings there)."""

        def __init__(self, mods: dict):
            self.mod = dict(mods)
            self.obsm: dict = {}
            first = next(iter(self.mod.values()), None)
            self.obs = pd.DataFrame(index=first.obs_names if first is not None and first.X is not None else pd.Index([]))

        @property
        def n_mod(self) -> int:
            return len(self.mod)

        @property
        def obs_names(self):
            return self.obs.index

        @property
        def n_obs(self) -> int:
            return len(self.obs.index)

        def __getitem__(self, name: str):
            return self.mod[name]

        def update(self) -> None:
            return

        def __repr__(self) -> str:
            return f"MuData object with n_obs = {self.n_obs}, modalities {list(self.mod)}"
