"""Import-path alias of ``rlaopt/preconditioners/factory.py``."""
from ._precond import _get_precond  # noqa: F401

__all__ = ["_get_precond"]
