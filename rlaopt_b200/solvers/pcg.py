"""Import-path alias of ``rlaopt/solvers/pcg.py``."""
from ._pcg import PCG  # noqa: F401
