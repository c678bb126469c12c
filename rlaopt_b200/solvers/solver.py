"""Import-path alias of ``rlaopt/solvers/solver.py``."""
from ._solver import Solver  # noqa: F401

__all__ = ["Solver"]
