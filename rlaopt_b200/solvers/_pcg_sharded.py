"""Block PCG with ROW-SHARDED solver state over a ``RowShardedLinOp`` (one process per GPU).

Same iteration as ``rlaopt/solvers/pcg.py:32-93`` (and ``_pcg.PCG``), but every rank keeps only its rows
``[lo, hi)`` of ``W``, ``R``, ``Z``, the search directions ``P`` and the Nystrom factor ``U``:

    per step   all-gather of the directions (the product needs every row of its input)       n x k
               local fused product  (A P)_loc + reg P_loc  with the Gram partial  P_loc^T (A P)_loc
               all-reduce of the k x k Gram partials, of  U_loc^T R_loc  (r x k) and of  R_loc^T Z_loc  (k x k)
    build      Y_loc = A_loc Omega (no gather of the n x r sketch), all-reduce of the r x r cores
               Omega_loc^T Y_loc and F_loc^T F_loc, replicated r x r Cholesky / eigh, U_loc = F_loc W

so the tall-skinny work (``U^T R``, ``U c``, the n x r factorisation) is divided by the number of ranks and the only
n-sized traffic is one all-gather of the directions per iteration.  Iterates agree with the replicated solver up to
the summation order of the reductions.  Supports the Nystrom and identity preconditioners (the ones the KRR configs
use); ``_get_solver`` falls back to the replicated ``PCG`` for anything else.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from rlaopt_b200.linops.fused import apply_fused
from rlaopt_b200.linops.spmd import RowShardedLinOp
from rlaopt_b200.preconditioners import IdentityConfig, NystromConfig, PreconditionerConfig
from rlaopt_b200.preconditioners._configs import _DampingMode
from rlaopt_b200.sketches import get_sketch
from rlaopt_b200.utils import rng as _rng

from ._pcg import _small_solve
from ._solver import Solver


class _ShardedNystrom:
    """Randomized Nystrom preconditioner ``U diag(S) U^T + rho I`` with the rows of ``U`` sharded over the ranks
    (construction of ``rlaopt/preconditioners/nystrom.py:55-98``, inverse of ``:112-132``)."""

    def __init__(self, config: NystromConfig, A: RowShardedLinOp, device: torch.device):
        self.config, self.A, self.group = config, A, A.group
        lo, hi = A.lo, A.hi
        dtype = A.dtype
        self.low_precision = dtype != torch.float64
        # Omega is drawn replicated (replicated_rng broadcasts it): the product needs all of its rows
        Omega = get_sketch(config.sketch, "right", config.rank, A.shape[1], dtype=dtype, device=device).Omega_mat
        if _rng._REPLICATED[0] is None and dist.get_world_size(self.group) > 1:
            # ranks seeded differently would sketch with different matrices: take rank 0's
            dist.broadcast(Omega, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        Y_loc = A.local_matmat(Omega)  # (n_loc, r): this rank's rows of A @ Omega
        core = Omega[lo:hi].T @ Y_loc  # partial of Omega^T Y
        dist.all_reduce(core, group=self.group)
        shift = torch.finfo(dtype).eps * torch.trace(core)
        core.diagonal().add_(shift)
        C = torch.linalg.cholesky(core, upper=False)
        F_loc = torch.linalg.solve_triangular(C.T, Y_loc, upper=True, left=False) if hi > lo else Y_loc
        r = config.rank
        if self.low_precision:
            # fp32 operators: singular vectors of the tall factor through its fp64 Gram matrix (as the replicated
            # build, preconditioners/_precond.py), the r x r partials all-reduced, eigh replicated
            G = F_loc.double().T @ F_loc.double()
            dist.all_reduce(G, group=self.group)
            evals, V = torch.linalg.eigh(G)
            evals, V = evals.flip(0), V.flip(1)
            sig2 = evals.clamp_min(0.0)
            inv_sig = torch.where(sig2 > 0, sig2.clamp_min(torch.finfo(torch.float64).tiny).rsqrt(), torch.zeros_like(sig2))
            self.U = (F_loc.double() @ (V * inv_sig)).to(dtype)  # (n_loc, r)
        else:
            # fp64 operators: tall-skinny QR over the ranks -- local QR, all-gather of the r x r triangles, QR of the
            # stack, SVD of the final triangle (the Gram route would square the condition number)
            world, rank_id = dist.get_world_size(self.group), dist.get_rank(self.group)
            Q1 = F_loc.new_zeros((F_loc.shape[0], r))
            R1 = F_loc.new_zeros((r, r))
            if hi > lo:
                q, t = torch.linalg.qr(F_loc, mode="reduced")
                Q1[:, : q.shape[1]], R1[: t.shape[0]] = q, t
            stack = F_loc.new_empty((world * r, r))
            dist.all_gather_into_tensor(stack, R1, group=self.group)
            Q2, R2 = torch.linalg.qr(stack, mode="reduced")
            Ur, sig, _ = torch.linalg.svd(R2, full_matrices=False)
            self.U = Q1 @ (Q2[rank_id * r:(rank_id + 1) * r] @ Ur)
            sig2 = sig * sig
        self.S = torch.clamp(sig2.to(dtype) - shift, min=0.0)
        self.L = None

    def update_damping(self, baseline_rho: float) -> None:
        if self.config.damping_mode == _DampingMode.ADAPTIVE:
            self.config.rho = baseline_rho + self.S[-1]
            self.L = None

    def solve_local(self, R_loc: torch.Tensor) -> torch.Tensor:
        """This rank's rows of ``P^{-1} R``."""
        rho = self.config.rho
        UTR = self.U.T @ R_loc
        dist.all_reduce(UTR, group=self.group)
        if self.low_precision:
            if self.L is None:
                G = self.U.T @ self.U
                dist.all_reduce(G, group=self.group)
                G.diagonal().add_(rho / self.S.clamp_min(torch.finfo(self.S.dtype).tiny))
                self.L = torch.linalg.cholesky(G)
            return (R_loc - self.U @ torch.cholesky_solve(UTR, self.L, upper=False)) / rho
        return (R_loc - self.U @ UTR) / rho + self.U @ (UTR / (self.S + rho).unsqueeze(-1))


class _ShardedIdentity:
    def update_damping(self, baseline_rho: float) -> None:
        pass

    def solve_local(self, R_loc: torch.Tensor) -> torch.Tensor:
        return R_loc


def sharded_pcg_supported(system, precond_config: PreconditionerConfig) -> bool:
    return (isinstance(system.A, RowShardedLinOp) and system.A.shape[0] == system.A.shape[1]
            and type(precond_config) in (NystromConfig, IdentityConfig))


class ShardedPCG(Solver):
    def __init__(self, system, W_init: torch.Tensor, precond_config: PreconditionerConfig, device: torch.device):
        A = system.A
        self.system, self.A, self.group = system, A, A.group
        self.device = device
        self.lo, self.hi = A.lo, A.hi
        lo, hi = self.lo, self.hi
        self.B_loc = system.B[lo:hi]
        self._W_loc = W_init[lo:hi].clone()
        self._W_full, self._W_stamp, self._stamp = None, -1, 0
        if type(precond_config) is NystromConfig:
            self.P = _ShardedNystrom(precond_config, A, device)
        else:
            self.P = _ShardedIdentity()
        self.P.update_damping(baseline_rho=system.reg)
        # R = B - (A + reg I) W on this rank's rows
        self.R = self._residual_local(W_init)
        self.Z = self.P.solve_local(self.R)
        self.P_ = self.Z.clone()
        self.RZ = self._reduce(self.R.T @ self.Z)

    # -- helpers -------------------------------------------------------------
    def _reduce(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, group=self.group)
        return t

    def _residual_local(self, W_full: torch.Tensor) -> torch.Tensor:
        lo, hi = self.lo, self.hi
        if self.A.local_op is None:
            return W_full.new_zeros((0, W_full.shape[1]))
        R, _, _ = apply_fused(self.A.local_op, W_full, alpha=-1.0, addend=W_full[lo:hi], beta=-self.system.reg,
                              rhs=self.B_loc, gamma=1.0)
        return R

    def _gather(self, T_loc: torch.Tensor) -> torch.Tensor:
        return self.A._gather_rows(T_loc)

    @property
    def W(self) -> torch.Tensor:
        """Full iterate (all-gathered on demand, once per iteration)."""
        if self._W_stamp != self._stamp:
            self._W_full, self._W_stamp = self._gather(self._W_loc), self._stamp
        return self._W_full

    def residual_sqnorms(self) -> torch.Tensor:
        """Squared column norms of the recurrence residual (``LinSys.solve(..., residual="recurrence")``)."""
        return self._reduce((self.R * self.R).sum(dim=0))

    def _restart_from_residual(self, R_full: torch.Tensor) -> None:
        self.R = R_full[self.lo:self.hi].clone()
        self.Z = self.P.solve_local(self.R)
        self.P_ = self.Z.clone()
        self.RZ = self._reduce(self.R.T @ self.Z)

    # -- one iteration ---------------------------------------------------------
    def _step(self):
        mask = self.system.mask
        if not bool(mask.any()):
            return
        all_active = bool(mask.all())
        idx = None if all_active else torch.nonzero(mask.to(self.R.device)).squeeze(-1)
        D_loc = self.P_ if all_active else self.P_[:, idx].contiguous()
        RZ = self.RZ if all_active else self.RZ[idx][:, idx]
        D_full = self._gather(D_loc)  # the product's input: every row of the directions
        k_act = D_loc.shape[1]
        if self.A.local_op is None:
            AD_loc, G = D_loc, D_loc.new_zeros((k_act, k_act))
        else:
            AD_loc, G, _ = apply_fused(self.A.local_op, D_full, addend=D_loc, beta=self.system.reg, gram_with=D_loc)
        G = self._reduce(G.contiguous())
        alpha = _small_solve(G, RZ)
        if all_active:
            self._W_loc.addmm_(D_loc, alpha)
            self.R.addmm_(AD_loc, alpha, alpha=-1.0)
            R_act = self.R
        else:
            self._W_loc[:, idx] += D_loc @ alpha
            R_act = self.R[:, idx] - AD_loc @ alpha
            self.R[:, idx] = R_act
        Z_act = self.P.solve_local(R_act)
        RZ_new = self._reduce(R_act.T @ Z_act)
        beta = _small_solve(RZ, RZ_new)
        if all_active:
            self.Z = Z_act
            self.P_ = torch.addmm(Z_act, D_loc, beta)
            self.RZ = RZ_new
        else:
            self.Z[:, idx] = Z_act
            self.P_[:, idx] = Z_act + D_loc @ beta
            full = torch.zeros_like(self.RZ)
            full[idx.unsqueeze(1), idx.unsqueeze(0)] = RZ_new
            self.RZ = full
        self._stamp += 1
