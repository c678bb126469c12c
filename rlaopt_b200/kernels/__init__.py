"""Implicit kernel-matrix linear operators (mirror of ``rlaopt.kernels``)."""
from . import configs, standard
from .configs import *  # noqa: F401,F403
from .standard import *  # noqa: F401,F403

__all__ = list(configs.__all__) + list(standard.__all__)
