"""Import-path alias of ``rlaopt/solvers/factory.py``."""
from ._solver import _get_solver  # noqa: F401

__all__ = ["_get_solver"]
