"""Random test matrices for the Nystrom sketch ``Y = K @ Omega`` (mirror of ``rlaopt.sketches``).

Public surface of the reference (``rlaopt/sketches/__init__.py``): ``get_sketch`` and the
``Sketch`` interface (``Omega_mat``, ``_apply_left/_apply_right/_apply_left_trans/
_apply_right_trans``, ``sketches/sketch.py:17-117``).  ``gauss`` and ``ortho`` are the sketches
on the kernel-matmat path (``NystromConfig.sketch`` defaults to ``"ortho"``,
``preconditioners/configs.py:81``); the sparse sign sketch belongs to the reference's CSC
kernels, which are outside this package's scope, and raises ``NotImplementedError``.
"""
from ._sketch import Gauss, Ortho, Sketch, get_sketch

__all__ = ["get_sketch", "Sketch", "Gauss", "Ortho"]
