"""Host-side parameter initialisation (mirrors reference initialization/)."""

from .initialize import initialize_mat, initialize_standard_nmf  # noqa: F401
