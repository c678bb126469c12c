"""Solver base class and the config -> solver factory (``solvers/solver.py``, ``solvers/factory.py``)."""
from __future__ import annotations

import torch


class Solver:
    """A solver owns the iterate ``W`` and advances it with ``_step()``."""

    def _get_precond(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _step(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError


def _get_solver(model, W_init: torch.Tensor, solver_config):
    from ._configs import PCGConfig, SAPConfig
    from ._pcg import PCG
    from ._sap import SAP

    if type(solver_config) is PCGConfig:
        import os

        from ._pcg_sharded import ShardedPCG, sharded_pcg_supported

        # one process per GPU over a row-sharded operator: keep W / R / Z / P and the Nystrom factor row-sharded too
        # (RLAOPT_B200_SHARDED_STATE=0 keeps the replicated solver)
        if os.environ.get("RLAOPT_B200_SHARDED_STATE", "1") != "0" and sharded_pcg_supported(model, solver_config.precond_config):
            return ShardedPCG(system=model, W_init=W_init, precond_config=solver_config.precond_config,
                              device=solver_config.device)
        return PCG(system=model, W_init=W_init, precond_config=solver_config.precond_config,
                   device=solver_config.device)
    if type(solver_config) is SAPConfig:
        return SAP(system=model, W_init=W_init, precond_config=solver_config.precond_config,
                   device=solver_config.device, blk_sz=solver_config.blk_sz, accel=solver_config.accel,
                   accel_config=solver_config.accel_config, power_iters=solver_config.power_iters)
    return None
