"""Model namespace, mirroring reference models/__init__.py:4-15."""

from .corrnmf import CorrNMFDet
from .klnmf import KLNMF
from .mmcorrnmf import MultimodalCorrNMF
from .mvnmf import MvNMF

__all__ = ["CorrNMFDet", "KLNMF", "MultimodalCorrNMF", "MvNMF"]
