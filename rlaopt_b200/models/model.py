"""Import-path alias of ``rlaopt/models/model.py``."""
from ._linsys import Model  # noqa: F401

__all__ = ["Model"]
