"""Iterative solvers whose cost is the kernel-matmat path (mirror of ``rlaopt.solvers``).

Same public names as the reference (``solvers/__init__.py``): ``SAPAccelConfig``,
``SolverConfig``, ``PCGConfig``, ``SAPConfig``, ``Solver`` and the private helpers
``_is_solver_config``, ``_get_solver_name``, ``_get_solver`` that ``LinSys.solve`` uses.
"""
from ._configs import (PCGConfig, SAPAccelConfig, SAPConfig, SolverConfig, _get_solver_name, _is_solver_config)
from ._pcg import PCG
from ._sap import SAP
from ._solver import Solver, _get_solver

__all__ = ["SAPAccelConfig", "SolverConfig", "PCGConfig", "SAPConfig", "_is_solver_config", "_get_solver_name",
           "_get_solver", "Solver"]
