"""Host-side helpers (validators) for the kernel-matmat path."""
from .input_checkers import *  # noqa: F401,F403
from . import input_checkers as _ic

__all__ = list(_ic.__all__)
