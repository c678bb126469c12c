"""Sketch-and-project (block coordinate descent) solver, with Nesterov acceleration = ASkotch.

Step of ``rlaopt/solvers/sap.py:129-175``: draw a row block ``blk`` on the host
(``torch.multinomial`` over uniform probabilities, ``sap.py:52-63``), build a preconditioner for
the block system ``A[blk, blk] + reg I`` from the block oracle, estimate the step size by power
iteration on ``P^{-1}(A_bb + reg I)`` (``sap.py:80-110``), form the block gradient
``A[blk, :] @ Y + reg Y[blk] - B[blk]`` with the row oracle and update ``W`` (and the momentum
sequences ``V``, ``Y``) on the rows of the block only.

The operator ``A_blk_oracle(blk)`` is built once per step and shared by the preconditioner and
all power iterations (the reference rebuilds it for every matvec), and the
block updates are scattered with ``index_add_`` instead of through dense zero matrices.

Block sampling.  ``torch.multinomial`` over n uniform probabilities on the CPU costs O(n) per step
(20 ms at n = 1M, 0.2 s at n = 10M -- more than the GPU work of the step).  Two measures:
the next block is drawn by a helper thread while the GPU works on the current step (from a private generator
forked off the global CPU stream when the solve starts, so callbacks and user code that draw random numbers never
interleave with it; under ``host_rng`` -- the parity setting -- the prefetch is off and the blocks come from the global
stream in the reference's order),
and ``SAP.block_sampler = "device"`` (or ``RLAOPT_B200_SAP_SAMPLER=device``) draws the block with
``torch.multinomial`` on the GPU instead (different stream than the reference, no host work at all).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from rlaopt_b200.linops import LinOp
from rlaopt_b200.linops.fused import apply_fused
from rlaopt_b200.preconditioners import (IdentityConfig, NewtonConfig, NystromConfig, Preconditioner,
                                         PreconditionerConfig, _get_precond)
from rlaopt_b200.spectral_estimators import randomized_powering
from rlaopt_b200.utils import host_rng_enabled, sync_from_rank0

from ._configs import SAPAccelConfig
from ._solver import Solver

VALID_PRECONDS = [IdentityConfig, NewtonConfig, NystromConfig]


class SAP(Solver):
    def __init__(self, system, W_init: torch.Tensor, precond_config: PreconditionerConfig, device: torch.device,
                 blk_sz: int, accel: bool, accel_config: SAPAccelConfig | None, power_iters: int):
        if type(precond_config) not in VALID_PRECONDS:
            raise TypeError(
                f"Valid preconditioner configs for SAP are {VALID_PRECONDS}, but received {type(precond_config)}")
        self.system = system
        self.precond_config = precond_config
        self.device = device
        self.blk_sz = blk_sz
        self.accel = accel
        self.accel_config = accel_config
        self.power_iters = power_iters
        self._W = W_init.clone()

        n = system.A.shape[0]
        self.probs = torch.ones(n) / n  # host tensor: blocks are sampled on the CPU, as in the reference
        self.probs_cpu = self.probs.numpy()
        self.block_sampler = os.environ.get("RLAOPT_B200_SAP_SAMPLER", "host")  # "host" (reference stream) | "device"
        self.prefetch_blocks = os.environ.get("RLAOPT_B200_SAP_PREFETCH", "1") != "0"
        self._probs_dev = None
        self._pool = None
        self._next_blk = None
        # Block draws on the helper thread use a PRIVATE generator forked from the global CPU stream at construction:
        # the global stream is then untouched by the prefetch (callbacks and user code that draw random numbers do not
        # interleave with it, and no extra block is consumed from it after the last step).  Without prefetch the
        # blocks come from the global stream itself, in the reference's order (solver parity runs under host_rng).
        self._gen = None
        self._np_rng = None
        if accel:
            mu, nu = accel_config.mu, accel_config.nu
            self.beta = 1 - (mu / nu) ** 0.5
            self.gamma = 1 / (mu * nu) ** 0.5
            self.alpha = 1 / (1 + self.gamma * nu)
            self.V = self._W.clone()
            self.Y = self._W.clone()

    @property
    def W(self):
        return self._W

    # ---- pieces of one step ----
    def _draw_host_blk(self) -> torch.Tensor:
        try:
            return torch.multinomial(self.probs, self.blk_sz, replacement=False, generator=self._gen)
        except RuntimeError as err:  # more than 2^24 categories
            if "number of categories cannot exceed" not in str(err):
                raise
            rng = np.random if self._gen is None else self._np_rng
            pick = rng.choice(self.probs.shape[0], size=self.blk_sz, replace=False, p=self.probs_cpu)
            return torch.from_numpy(pick)

    def _get_blk(self) -> torch.Tensor:
        if self.block_sampler == "device" and torch.device(self.device).type == "cuda":
            if self._probs_dev is None:
                self._probs_dev = self.probs.to(self.device)
            if self._probs_dev.numel() < (1 << 24):
                blk = torch.multinomial(self._probs_dev, self.blk_sz, replacement=False)
            else:  # uniform sampling without replacement
                blk = torch.randperm(self._probs_dev.numel(), device=self.device)[: self.blk_sz]
            return sync_from_rank0(blk, self.device)
        prefetch = self.prefetch_blocks and torch.device(self.device).type == "cuda" and not host_rng_enabled()
        if not prefetch:
            blk = self._draw_host_blk()
        else:  # drawing one step ahead on a helper thread, from the solver's private stream
            if self._pool is None:
                self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="sap-blocks")
                seed = int(torch.randint(0, 2**62, (1,)).item())  # one draw from the global stream seeds the fork
                self._gen = torch.Generator().manual_seed(seed)
                self._np_rng = np.random.default_rng(seed)
            blk = self._next_blk.result() if self._next_blk is not None else self._draw_host_blk()
            self._next_blk = self._pool.submit(self._draw_host_blk)
        return sync_from_rank0(blk, self.device)  # SPMD runs: every rank works on rank 0's block

    def close(self) -> None:
        """Stop the block-prefetch thread (idempotent; also runs when the solver is collected)."""
        pool, self._pool, self._next_blk = self._pool, None, None
        if pool is not None:
            pool.shutdown(wait=False, cancel_futures=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _get_precond(self, blk: torch.Tensor, A_bb=None) -> Preconditioner:
        P = _get_precond(self.precond_config)
        P._update(A_bb if A_bb is not None else self.system.A_blk_oracle(blk), self.device)
        P._update_damping(baseline_rho=self.system.reg)
        return P

    def _get_stepsize(self, blk: torch.Tensor, blk_precond: Preconditioner, A_bb=None):
        reg = self.system.reg
        if isinstance(self.precond_config, NewtonConfig):
            if self.precond_config.rho == reg:
                return 1.0  # the block preconditioner is the exact block inverse
            raise ValueError("SAP with a Newton preconditioner needs rho == reg (sap.py:89-93 leaves the step "
                             "size undefined otherwise)")
        if A_bb is None:
            A_bb = self.system.A_blk_oracle(blk)

        # P^{-1}(A_bb + reg I) is similar to the symmetric P^{-1/2}(A_bb + reg I)P^{-1/2}: same top eigenvalue
        S = LinOp(device=self.device, shape=torch.Size((self.blk_sz, self.blk_sz)),
                  matvec=blk_precond._inverse_matmul_compose(lambda v: A_bb @ v + reg * v),
                  dtype=self._W.dtype)
        max_eig, _ = randomized_powering(S, max_iters=self.power_iters)
        return max_eig ** (-1.0)

    def _get_block_update(self, W: torch.Tensor, B: torch.Tensor, blk: torch.Tensor, blk_precond: Preconditioner):
        # A[blk, :] W + reg W[blk] - B[blk] in the row oracle's output stage (sap.py:113-127)
        grad, _, _ = apply_fused(self.system.A_row_oracle(blk), W, addend=W, beta=self.system.reg, addend_idx=blk,
                                 rhs=B, gamma=-1.0, rhs_idx=blk)
        return blk_precond._inv @ grad

    def _step(self):
        mask = self.system.mask
        if not bool(mask.any()):
            return
        blk = self._get_blk()
        A_bb = self.system.A_blk_oracle(blk)
        blk_precond = self._get_precond(blk, A_bb)
        step = self._get_stepsize(blk, blk_precond, A_bb)

        dev = self._W.device
        rows = blk.to(dev)
        all_active = bool(mask.all())
        cols = None if all_active else torch.nonzero(mask.to(dev)).squeeze(-1)

        def take(T):
            return T if all_active else T[:, cols]

        src = self.Y if self.accel else self._W
        direction = self._get_block_update(take(src), take(self.system.B), rows, blk_precond)
        upd = step * direction  # (blk_sz, active columns)

        if not self.accel:
            if all_active:
                self._W.index_add_(0, rows, upd, alpha=-1.0)
            else:
                self._W[rows.unsqueeze(1), cols.unsqueeze(0)] -= upd
            return

        beta, gamma, alpha = self.beta, self.gamma, self.alpha
        if all_active:
            W, V, Y = self._W, self.V, self.Y
            W.copy_(Y).index_add_(0, rows, upd, alpha=-1.0)
            V.mul_(beta).add_(Y, alpha=1 - beta).index_add_(0, rows, upd, alpha=-gamma)
            torch.add(alpha * V, W, alpha=1 - alpha, out=Y)
        else:
            Wa = self.Y[:, cols].clone()
            Wa[rows] -= upd
            Va = beta * self.V[:, cols] + (1 - beta) * self.Y[:, cols]
            Va[rows] -= gamma * upd
            self._W[:, cols] = Wa
            self.V[:, cols] = Va
            self.Y[:, cols] = alpha * Va + (1 - alpha) * Wa
