"""Preconditioners that consume the kernel-matmat path (mirror of ``rlaopt.preconditioners``).

Same public names as the reference (``preconditioners/__init__.py``): the config dataclasses,
``Preconditioner`` and the ``_get_precond`` factory.  ``Identity``, ``Newton`` and ``Nystrom`` are
the preconditioners PCG and SAP/ASkotch build on kernel operators (``solvers/sap.py:23``);
``SkPreConfig`` is accepted for API compatibility but its sparse-sketch preconditioner depends on
the reference's CSC kernels (out of scope) and the factory raises ``NotImplementedError`` for it.
"""
from ._configs import (IdentityConfig, NewtonConfig, NystromConfig, PreconditionerConfig, SkPreConfig,
                       _is_precond_config)
from ._precond import Identity, Newton, Nystrom, Preconditioner, _get_precond

__all__ = ["PreconditionerConfig", "IdentityConfig", "NewtonConfig", "NystromConfig", "SkPreConfig",
           "_is_precond_config", "_get_precond", "Preconditioner"]
