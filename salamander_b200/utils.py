"""
Host-side argument checkers with the reference's names and error behaviour
(reference utils.py:16-99): TypeError for a wrong type, ValueError for a bad value,
shape or dictionary key.  Validation only -- no arithmetic of the fitting path lives here.
"""

from __future__ import annotations

from typing import Any, Iterable

import numpy as np
import pandas as pd

EPSILON = float(np.finfo(np.float32).eps)


def type_checker(arg_name: str, arg: Any, allowed_types) -> None:
    """TypeError unless ``type(arg)`` is exactly one of ``allowed_types`` (reference utils.py:59-77)."""
    allowed = [allowed_types] if isinstance(allowed_types, type) else list(allowed_types)
    if type(arg) not in allowed:
        raise TypeError(f"The type of '{arg_name}' has to be one of {allowed}.")


def value_checker(arg_name: str, arg: Any, allowed_values: Iterable[Any]) -> None:
    """ValueError unless ``arg`` is one of ``allowed_values`` (reference utils.py:80-99)."""
    if arg not in allowed_values:
        raise ValueError(f"The value of '{arg_name}' has to be one of {allowed_values}.")


def shape_checker(arg_name: str, arg, allowed_shape: tuple[int, ...]) -> None:
    """ValueError unless the array / frame has exactly ``allowed_shape`` (reference utils.py:38-56)."""
    type_checker(arg_name, arg, [np.ndarray, pd.DataFrame])
    if tuple(arg.shape) != tuple(allowed_shape):
        raise ValueError(f"The shape of '{arg_name}' has to be {allowed_shape}.")


def dict_checker(dict_name: str, dictionary: dict, valid_keys: list) -> None:
    """ValueError if the dictionary has a key outside ``valid_keys`` (reference utils.py:16-35)."""
    type_checker(dict_name, dictionary, dict)
    for key in dictionary:
        if key not in valid_keys:
            raise ValueError(f"'{dict_name}' includes keys outside of {valid_keys}.")


def normalize_WH(W: np.ndarray, H: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Host version used once by initialisation: (W / colsum, H * colsum) (reference utils.py:155-158).

    The per-iteration use inside MvNMF's line search runs on the device
    (sal_mvnmf_trial + the h_scale argument of sal_klnmf_pass).
    """
    # the reference's normalize_WH is numba-compiled: its np.sum(axis=0) adds the rows one after the other, whereas numpy
    # sums pairwise -- the last bit can differ.  Same order here, so a seed gives the reference's W0, H0 bit for bit.
    s = column_sums_sequential(W)
    return W / s, H * s[:, None]


def column_sums_sequential(W: np.ndarray) -> np.ndarray:
    """``sum_v W[v, j]`` with the rows added one after the other (the order of numba's ``np.sum(W, axis=0)``)."""
    s = np.zeros(W.shape[1], dtype=np.result_type(W.dtype, np.float64))
    for row in np.asarray(W):
        s += row
    return s


def match_to_catalog(signatures: pd.DataFrame, catalog: pd.DataFrame, metric: str = "cosine") -> pd.DataFrame:
    """For every signature the closest catalog entry under ``metric`` (reference utils.py:161-170)."""
    from sklearn.metrics import pairwise_distances

    similarity = 1 - pairwise_distances(signatures, catalog, metric=metric)
    return catalog.iloc[[int(np.argmax(row)) for row in similarity]]


def match_signatures_pair(signatures1: pd.DataFrame, signatures2: pd.DataFrame, metric: str = "cosine") -> np.ndarray:
    """Indices that reorder ``signatures2`` so that the summed pairwise distance to ``signatures1`` is minimal
    (assignment problem; reference utils.py:173-192).  Used to align restarts and models of a k-sweep."""
    from scipy.optimize import linear_sum_assignment
    from sklearn.metrics import pairwise_distances

    if signatures1.shape != signatures2.shape:
        raise ValueError("The signatures must be of the same shape.")
    return linear_sum_assignment(pairwise_distances(signatures1, signatures2, metric=metric))[1]
