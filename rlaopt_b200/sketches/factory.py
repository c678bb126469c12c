"""Import-path alias of ``rlaopt/sketches/factory.py``."""
from ._sketch import get_sketch  # noqa: F401

__all__ = ["get_sketch"]
