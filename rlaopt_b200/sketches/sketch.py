"""Import-path alias of ``rlaopt/sketches/sketch.py``."""
from ._sketch import Sketch  # noqa: F401
