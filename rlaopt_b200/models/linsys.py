"""Import-path alias of ``rlaopt/models/linsys.py``."""
from ._linsys import LinSys  # noqa: F401

__all__ = ["LinSys"]
