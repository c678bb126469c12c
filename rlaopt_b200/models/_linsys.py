"""``LinSys``: positive-definite systems ``(A + reg I) W = B`` solved by PCG or SAP/ASkotch.

Interface and control flow of ``rlaopt/models/model.py:14-128`` and ``models/linsys.py:14-159``:
``solve(solver_config, W_init, callback_fn, ..., callback_freq)`` logs every ``callback_freq``
iterations (each log evaluates the true residual with one full matmat, ``linsys.py:96-99``),
freezes converged right-hand sides through ``mask`` (``linsys.py:101-107``) and stops when all
have converged -- so iteration counts are multiples of ``callback_freq`` exactly as in the
reference.  The mask is kept on the host (it is tiny and only steers control flow).
"""
from __future__ import annotations

from typing import Any, Callable
from warnings import warn

import torch

from rlaopt_b200.linops.fused import apply_fused
from rlaopt_b200.linops.types import _is_linop_or_torch_tensor
from rlaopt_b200.utils import Logger, _is_callable, _is_nonneg_float, _is_torch_tensor


class Model:
    """Shared training loop: ``solver._step()`` + periodic logging / termination."""

    def _check_inputs(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _compute_internal_metrics(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _check_termination_criteria(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _get_log_fn(self, callback_fn: Callable | None, callback_args: list, callback_kwargs: dict):
        def log_fn(w):
            out = {}
            if callback_fn is not None:
                out["callback"] = callback_fn(w, self, *callback_args, **callback_kwargs)
            out["internal_metrics"] = self._compute_internal_metrics(w)
            return out

        return log_fn

    def _get_wandb_kwargs(self, log_in_wandb: bool, wandb_init_kwargs: dict | None, solver_name: str, solver_config,
                          callback_freq: int):
        if not log_in_wandb:
            return None
        kwargs = {"config": {"solver_name": solver_name, "solver_config": solver_config.to_dict(),
                             "callback_freq": callback_freq}}
        for key, value in (wandb_init_kwargs or {}).items():
            if key == "config":
                warn("Found 'config' key in wandb_init_kwargs. Merging with internally specified 'config' key.")
                kwargs["config"].update(value)
            else:
                kwargs[key] = value
        return kwargs

    def _train(self, logger: Logger, termination_fn: Callable, solver, max_iters: int):
        self._solver = solver  # the metrics may read the solver's own residual (LinSys, residual="recurrence")
        log = {0: logger._compute_log(0, solver.W)}
        if termination_fn(log[0]["metrics"]["internal_metrics"]):
            return solver.W, log
        for i in range(1, max_iters + 1):
            solver._step()
            # the iterate is only needed on logged iterations (a sharded solver assembles it on demand)
            entry = logger._compute_log(i, solver.W) if i % logger.log_freq == 0 else None
            if entry is not None:
                log[i] = entry
                if termination_fn(entry["metrics"]["internal_metrics"]):
                    break
        logger._terminate()
        close = getattr(solver, "close", None)
        if close is not None:
            close()  # e.g. the block-prefetch thread of SAP
        return solver.W, log


class LinSys(Model):
    def __init__(self, A, B: torch.Tensor, reg: float = 0.0, A_row_oracle: Callable | None = None,
                 A_blk_oracle: Callable | None = None):
        self._check_inputs(A, B, reg, A_row_oracle, A_blk_oracle)
        self._A = A
        self._B = B.unsqueeze(-1) if B.ndim == 1 else B
        self._reg = reg
        self._A_row_oracle = A_row_oracle
        self._A_blk_oracle = A_blk_oracle
        self._mask = torch.ones(self._B.shape[1], dtype=torch.bool)
        self._B_norm = None
        self._residual_mode = "true"
        self._metrics_from_recurrence = False
        self._failed_confirmations = 0

    A = property(lambda self: self._A)
    B = property(lambda self: self._B)
    reg = property(lambda self: self._reg)
    A_row_oracle = property(lambda self: self._A_row_oracle)
    A_blk_oracle = property(lambda self: self._A_blk_oracle)
    mask = property(lambda self: self._mask)

    def _check_inputs(self, A: Any, B: Any, reg: Any, A_row_oracle: Any, A_blk_oracle: Any):
        _is_linop_or_torch_tensor(A, "A")
        _is_torch_tensor(B, "B")
        _is_nonneg_float(reg, "reg")
        for fn, name in ((A_row_oracle, "A_row_oracle"), (A_blk_oracle, "A_blk_oracle")):
            if fn is not None:
                _is_callable(fn, name)
        if (A_row_oracle is None) != (A_blk_oracle is None):
            missing, given = (("A_blk_oracle", "A_row_oracle") if A_blk_oracle is None
                              else ("A_row_oracle", "A_blk_oracle"))
            raise ValueError(f"{missing} must be provided if {given} is provided")

    def _rhs_norms(self) -> torch.Tensor:
        if self._B_norm is None:
            self._B_norm = torch.linalg.norm(self._B, dim=0, ord=2)
        return self._B_norm

    def _true_sq_residual(self, W: torch.Tensor) -> torch.Tensor:
        """Squared column norms of ``B - (A W + reg W)``: one fused pass, the residual itself is never stored."""
        _, _, sqn = apply_fused(self.A, W, alpha=-1.0, addend=W, beta=-self.reg, rhs=self.B, gamma=1.0,
                                want_sqnorm=True, store=False)
        return sqn

    def _compute_internal_metrics(self, W: torch.Tensor):
        """``{"abs_res", "rel_res"}`` per right-hand side (``linsys.py:96-99``).

        ``residual="true"`` (default, the reference's behaviour): one full product per logged iteration, with the
        subtraction and the column norms folded into the product's output stage.  ``residual="recurrence"`` (opt-in,
        block PCG only): the norms of the residual the solver already carries, ``R_t = R_{t-1} - (A P + reg P) alpha``
        (``pcg.py:64-67``) -- a logged iteration then costs one product instead of two.  The recurrence drifts from
        the true residual by rounding, so convergence is *confirmed* once with the true residual before the solve
        stops; if the confirmation fails the solve continues on true residuals."""
        solver = getattr(self, "_solver", None)
        sqnorms = getattr(solver, "residual_sqnorms", None)
        self._metrics_from_recurrence = False
        if self._residual_mode == "recurrence" and sqnorms is not None and solver.W is W:
            abs_res = sqnorms().sqrt()
            self._metrics_from_recurrence = True
        else:
            abs_res = self._true_sq_residual(W).sqrt()
        return {"abs_res": abs_res, "rel_res": abs_res / self._rhs_norms()}

    def _check_termination_criteria(self, internal_metrics: dict, atol: float, rtol: float):
        tol = torch.clamp(rtol * self._rhs_norms(), min=atol)
        self._mask = (internal_metrics["abs_res"] > tol).cpu()
        done = not bool(self._mask.any())
        if done and self._metrics_from_recurrence:
            # confirm with the true residual (one product, once per solve when the recurrence is accurate)
            R_true, _, sqn = apply_fused(self.A, self._solver.W, alpha=-1.0, addend=self._solver.W, beta=-self.reg,
                                         rhs=self.B, gamma=1.0, want_sqnorm=True)
            abs_res = sqn.sqrt() if sqn is not None else torch.linalg.norm(R_true, dim=0, ord=2)
            internal_metrics["abs_res"], internal_metrics["rel_res"] = abs_res, abs_res / self._rhs_norms()
            self._mask = (abs_res > tol).cpu()
            done = not bool(self._mask.any())
            if not done:
                # the recurrence had drifted: restart the solver from the true residual (columns that were frozen
                # carry no search direction any more) and keep going; after three failed confirmations every
                # logged iteration evaluates the true residual again
                self._failed_confirmations += 1
                restart = getattr(self._solver, "_restart_from_residual", None)
                if restart is not None:
                    restart(R_true)
                if restart is None or self._failed_confirmations >= 3:
                    self._residual_mode = "true"
        return done

    def solve(self, solver_config, W_init, callback_fn=None, callback_args=[], callback_kwargs={},
              callback_freq=10, log_in_wandb=False, wandb_init_kwargs=None, *, residual: str | None = None):
        """``residual`` (extension, keyword only): ``"true"`` -- the reference's metric, one extra product per logged
        iteration -- or ``"recurrence"``, see :meth:`_compute_internal_metrics`; default from
        ``RLAOPT_B200_RESIDUAL`` (``"true"``)."""
        import os

        from rlaopt_b200.solvers import _get_solver, _get_solver_name, _is_solver_config

        mode = residual if residual is not None else os.environ.get("RLAOPT_B200_RESIDUAL", "true")
        if mode not in ("true", "recurrence"):
            raise ValueError(f"residual must be 'true' or 'recurrence', got {mode!r}")
        self._residual_mode = mode
        self._failed_confirmations = 0

        _is_solver_config(solver_config, "solver_config")
        _is_torch_tensor(W_init, "W_init")
        if log_in_wandb and wandb_init_kwargs is None:
            raise ValueError("wandb_init_kwargs must be specified if log_in_wandb is True")
        atol, rtol = solver_config.atol, solver_config.rtol
        logger = Logger(
            log_freq=callback_freq,
            log_fn=self._get_log_fn(callback_fn, callback_args, callback_kwargs),
            wandb_kwargs=self._get_wandb_kwargs(log_in_wandb, wandb_init_kwargs, _get_solver_name(solver_config),
                                                solver_config, callback_freq),
        )
        solver = _get_solver(model=self, W_init=W_init, solver_config=solver_config)
        return self._train(logger=logger,
                           termination_fn=lambda m: self._check_termination_criteria(m, atol, rtol),
                           solver=solver, max_iters=solver_config.max_iters)
