"""Problem classes driving the solvers (mirror of ``rlaopt.models``): ``Model`` and ``LinSys``."""
from ._linsys import LinSys, Model

__all__ = ["Model", "LinSys"]
