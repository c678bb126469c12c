"""Block preconditioned conjugate gradient for ``(A + reg I) W = B`` with k right-hand sides.

Iteration of ``rlaopt/solvers/pcg.py:32-93``: the k search directions are coupled through the
k x k Gram systems ``alpha = (P^T A P)^{-1} (R^T Z)`` and ``beta = (R^T Z)^{-1} (R_new^T Z_new)``;
columns whose residual has converged (``system.mask``) are frozen.  One step costs one fused
kernel matmat ``A @ P`` with the active columns; everything else is tall-skinny BLAS on the
device.  When every column is active (the common case) the state is updated in place with
``addmm_`` instead of gathering / scattering masked copies.
"""
from __future__ import annotations

import torch

from rlaopt_b200.linops.fused import apply_fused
from rlaopt_b200.preconditioners import PreconditionerConfig, _get_precond

from ._solver import Solver


def _small_solve(G: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """``G^{-1} rhs`` for the k x k Gram systems; a single right-hand side is a division (no LU, no host sync)."""
    if G.shape[0] == 1:
        return rhs / G
    return torch.linalg.solve(G, rhs)


class PCG(Solver):
    def __init__(self, system, W_init: torch.Tensor, precond_config: PreconditionerConfig, device: torch.device):
        self.system = system
        self.precond_config = precond_config
        self.device = device
        self._W = W_init.clone()
        self.P = self._get_precond()
        # R = B - (A + reg I) W in the operator's output stage,  Z = P^{-1} R,  first directions = Z,  RZ = R^T Z
        self.R, _, _ = apply_fused(system.A, self._W, alpha=-1.0, addend=self._W, beta=-system.reg, rhs=system.B, gamma=1.0)
        self.Z = self.P._inv @ self.R
        self.P_ = self.Z.clone()
        self.RZ = self.R.T @ self.Z

    @property
    def W(self):
        return self._W

    def _apply(self, X: torch.Tensor, gram: bool = False):
        """``(A + reg I) X`` and, on request, ``X^T (A + reg I) X`` from the same pass (``pcg.py:58-61``)."""
        Y, G, _ = apply_fused(self.system.A, X, addend=X, beta=self.system.reg, gram_with=X if gram else None)
        return (Y, G) if gram else Y

    def residual_sqnorms(self) -> torch.Tensor:
        """Squared column norms of the recurrence residual (``LinSys.solve(..., residual="recurrence")``)."""
        return (self.R * self.R).sum(dim=0)

    def _restart_from_residual(self, R: torch.Tensor) -> None:
        """Restart block CG at the current iterate from a freshly evaluated residual (``LinSys`` calls this when the
        true residual contradicts the recurrence): new preconditioned residual, steepest-descent directions, Gram."""
        self.R = R.clone() if R.data_ptr() == self.R.data_ptr() else R
        self.Z = self.P._inv @ self.R
        self.P_ = self.Z.clone()
        self.RZ = self.R.T @ self.Z

    def _get_precond(self):
        P = _get_precond(self.precond_config)
        P._update(self.system.A, self.device)
        P._update_damping(baseline_rho=self.system.reg)
        return P

    def _step(self):
        mask = self.system.mask
        if not bool(mask.any()):
            return
        if bool(mask.all()):
            self._step_all()
        else:
            self._step_masked(mask.to(self._W.device))

    def _step_all(self):
        D = self.P_
        AD, G = self._apply(D, gram=True)
        alpha = _small_solve(G, self.RZ)
        self._W.addmm_(D, alpha)
        self.R.addmm_(AD, alpha, alpha=-1.0)
        self.Z = self.P._inv @ self.R
        RZ_new = self.R.T @ self.Z
        beta = _small_solve(self.RZ, RZ_new)
        self.P_ = torch.addmm(self.Z, D, beta)
        self.RZ = RZ_new

    def _step_masked(self, mask: torch.Tensor):
        idx = torch.nonzero(mask).squeeze(-1)
        D = self.P_[:, idx]
        RZ = self.RZ[idx][:, idx]
        AD, G = self._apply(D.contiguous(), gram=True)
        alpha = _small_solve(G, RZ)
        self._W[:, idx] += D @ alpha
        R_act = self.R[:, idx] - AD @ alpha
        self.R[:, idx] = R_act
        Z_act = self.P._inv @ R_act
        self.Z[:, idx] = Z_act
        RZ_new = R_act.T @ Z_act
        beta = _small_solve(RZ, RZ_new)
        self.P_[:, idx] = Z_act + D @ beta
        full = torch.zeros_like(self.RZ)
        full[idx.unsqueeze(1), idx.unsqueeze(0)] = RZ_new
        self.RZ = full
