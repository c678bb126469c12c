"""Import-path alias of ``rlaopt/preconditioners/newton.py``."""
from ._precond import Newton  # noqa: F401
