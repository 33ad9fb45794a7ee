"""Model namespace, mirroring reference models/__init__.py:4-15."""

from .klnmf import KLNMF
from .mvnmf import MvNMF

__all__ = ["KLNMF", "MvNMF"]
