"""Import-path alias of ``rlaopt/spectral_estimators/spectral_norm.py``."""
from . import randomized_powering  # noqa: F401

__all__ = ["randomized_powering"]
