import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference mounted at /root/reference")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible."""
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_path(*parts) -> str:
    return os.path.join(GOLDEN, *parts)


def load_counts(*parts) -> np.ndarray:
    """counts.csv fixtures are features x samples with an index column."""
    return pd.read_csv(golden_path(*parts), index_col=0).values


@pytest.fixture(scope="session")
def pcawg_sbs() -> np.ndarray:
    """PCAWG breast SBS counts as (V=96, D=192) int64."""
    path = os.path.join(ROOT, "salamander_b200", "data", "pcawg_breast_sbs.csv")
    return pd.read_csv(path, index_col=0).values
