"""
Oracle support (test infrastructure): import the LIVE reference when it is mounted.

``/root/reference`` exists only in the build container (never on the GPU box), so
everything here is optional: ``available()`` says whether it can be used.  The
reference needs anndata / mudata / matplotlib / seaborn / ... which are not
installed; they are replaced with stubs in ``sys.modules`` and AnnData by a small
stand-in (RangeIndex ``.obs`` so the positional ``obs["scalings"][k]`` look-ups at
reference models/corrnmf_det.py:108,135 keep working under pandas 3).

Used by ``oracle/make_golden.py`` (to generate tests/golden/trajectories/*) and by
the ``-m "not gpu"`` tests that assert restatement == live reference.
"""

from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get("SALAMANDER_REFERENCE", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "src")


def available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "salamander", "models", "_utils_klnmf.py"))


class RefAnnData:
    """Just enough of anndata.AnnData for the reference's fit paths."""

    def __init__(self, X=None, obs=None):
        if isinstance(X, pd.DataFrame):
            self.obs_names = pd.Index(X.index.astype(str))
            self.var_names = pd.Index(X.columns.astype(str))
            X = X.values
        elif X is not None:
            X = np.asarray(X)
            self.obs_names = pd.Index([str(i) for i in range(X.shape[0])])
            self.var_names = pd.Index([str(i) for i in range(X.shape[1])])
        else:
            self.obs_names = pd.Index([])
            self.var_names = pd.Index([])
        self.X = X
        n = 0 if X is None else X.shape[0]
        self.obs = pd.DataFrame(index=pd.RangeIndex(n))
        self.obsm = {}
        self.obsp = {}

    @property
    def n_obs(self):
        return 0 if self.X is None else self.X.shape[0]

    @property
    def n_vars(self):
        return 0 if self.X is None else self.X.shape[1]

    @property
    def shape(self):
        return (self.n_obs, self.n_vars)

    def copy(self):
        new = RefAnnData(None if self.X is None else np.array(self.X))
        new.obs_names = self.obs_names.copy()
        new.var_names = self.var_names.copy()
        new.obs = self.obs.copy()
        new.obsm = {k: np.array(v) for k, v in self.obsm.items()}
        new.obsp = {k: np.array(v) for k, v in self.obsp.items()}
        return new

    def to_df(self):
        return pd.DataFrame(self.X, index=self.obs_names, columns=self.var_names)

    def __getitem__(self, idx):
        rows, cols = idx if isinstance(idx, tuple) else (idx, slice(None))
        new = RefAnnData(np.asarray(self.X)[rows][:, cols])
        new.obs_names = self.obs_names[rows]
        new.var_names = self.var_names[cols]
        obs = self.obs.iloc[rows].reset_index(drop=True)
        new.obs = obs
        new.obsm = {k: np.asarray(v)[rows] for k, v in self.obsm.items()}
        return new


def _concat(adatas, join="outer"):
    out = RefAnnData(np.concatenate([np.asarray(a.X) for a in adatas], axis=0))
    out.obs_names = pd.Index(np.concatenate([np.asarray(a.obs_names) for a in adatas]))
    out.var_names = adatas[0].var_names
    return out


class RefMuData:
    """The part of ``mudata.MuData`` the reference's MultimodalCorrNMF touches: ``mod`` (name -> AnnData), ``n_mod``,
    ``n_obs``, ``obs_names``, ``obs``, ``obsm``, ``obsp``, ``update()`` and item access by modality name."""

    def __init__(self, mods):
        import pandas as pd

        self.mod = dict(mods)
        first = next(iter(self.mod.values()))
        self.obs_names = first.obs_names
        self.obs = pd.DataFrame(index=self.obs_names)
        self.obsm, self.obsp = {}, {}

    @property
    def n_mod(self):
        return len(self.mod)

    @property
    def n_obs(self):
        return len(self.obs_names)

    def update(self):
        return None

    def __getitem__(self, name):
        return self.mod[name]


_loaded = {}


def load_utils_klnmf():
    """The reference's models/_utils_klnmf.py as a stand-alone module (no package imports)."""
    if "utils_klnmf" not in _loaded:
        path = os.path.join(_SRC, "salamander", "models", "_utils_klnmf.py")
        spec = importlib.util.spec_from_file_location("_ref_utils_klnmf", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _loaded["utils_klnmf"] = mod
    return _loaded["utils_klnmf"]


def load_package():
    """``import salamander`` from the mounted reference with third-party stubs."""
    if "pkg" in _loaded:
        return _loaded["pkg"]
    ad = types.ModuleType("anndata")
    ad.AnnData = RefAnnData
    ad.concat = _concat
    sys.modules.setdefault("anndata", ad)
    for name in (
        "mudata",
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.axes",
        "matplotlib.colors",
        "matplotlib.patches",
        "matplotlib.lines",
        "matplotlib.figure",
        "seaborn",
        "fastcluster",
        "adjustText",
        "umap",
    ):
        sys.modules.setdefault(name, MagicMock())
    if _SRC not in sys.path:
        sys.path.insert(0, _SRC)
    sys.modules["mudata"].MuData = RefMuData
    pkg = importlib.import_module("salamander")
    _loaded["pkg"] = pkg
    return pkg
