"""Import-path alias of ``rlaopt/preconditioners/enums.py``."""
from ._configs import _DampingMode  # noqa: F401
