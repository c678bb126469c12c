"""Import-path alias of ``rlaopt/preconditioners/identity.py``."""
from ._precond import Identity  # noqa: F401
