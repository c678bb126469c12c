"""Import-path alias of ``rlaopt/sketches/ortho.py``."""
from ._sketch import Ortho  # noqa: F401
