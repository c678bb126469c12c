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
        # R = B - (A + reg I) W,  Z = P^{-1} R,  first directions = Z,  RZ = R^T Z
        self.R = system.B - self._apply(self._W)
        self.Z = self.P._inv @ self.R
        self.P_ = self.Z.clone()
        self.RZ = self.R.T @ self.Z

    @property
    def W(self):
        return self._W

    def _apply(self, X: torch.Tensor) -> torch.Tensor:
        """(A + reg I) X"""
        Y = self.system.A @ X
        return Y.add_(X, alpha=self.system.reg) if Y.data_ptr() != X.data_ptr() else Y + self.system.reg * X

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
        AD = self._apply(D)
        alpha = _small_solve(D.T @ AD, self.RZ)
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
        AD = self._apply(D)
        alpha = _small_solve(D.T @ AD, RZ)
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
