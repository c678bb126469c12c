"""Import-path alias of ``rlaopt/preconditioners/nystrom.py``."""
from ._precond import Nystrom  # noqa: F401
