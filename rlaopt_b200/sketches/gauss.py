"""Import-path alias of ``rlaopt/sketches/gauss.py``."""
from ._sketch import Gauss  # noqa: F401
