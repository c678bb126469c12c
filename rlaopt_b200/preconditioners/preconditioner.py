"""Import-path alias of ``rlaopt/preconditioners/preconditioner.py``."""
from ._precond import Preconditioner, _InvPreconditioner  # noqa: F401

__all__ = ["Preconditioner"]
