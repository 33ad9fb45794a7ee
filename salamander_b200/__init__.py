"""
salamander_b200 -- B200-native (sm_100a) device backend for the NMF fitting hot path of
parklab/Salamander: ``sal.models.KLNMF / MvNMF / CorrNMFDet / MultimodalCorrNMF`` with the
reference's constructor and ``fit(adata)`` API (namespace shape of reference __init__.py:6-19).

Host code is Python + PyTorch (device memory, streams, torch.distributed); every numerical
step is hand-written CUDA behind the C ABI in include/salamander_b200.h.
"""

from . import models, sweep
from ._anndata import AnnData, MuData

__version__ = "0.1.0"
__all__ = ["models", "sweep", "AnnData", "MuData"]
