"""Identity, Newton (exact Cholesky) and randomized Nystrom preconditioners.

Interface of ``rlaopt/preconditioners/preconditioner.py:18-180``: ``_update(A, device)``,
``P @ x``, ``P._inv @ x``, ``_inverse_matmul_compose(fn)``, ``_update_damping(baseline_rho)``.
All state lives on ``device``; the only O(n^2) work is the sketch ``Y = A @ Omega`` (Nystrom) or
``A @ I`` (Newton on a block operator), i.e. one fused kernel matmat.
"""
from __future__ import annotations

from typing import Callable

import torch

from rlaopt_b200.sketches import get_sketch
from rlaopt_b200.utils import _is_torch_tensor_1d_2d

from ._configs import (IdentityConfig, NewtonConfig, NystromConfig, PreconditionerConfig, SkPreConfig,
                       _DampingMode)


class _InvPreconditioner:
    """``P._inv @ x`` view of a preconditioner (``preconditioner.py:153-180``)."""

    def __init__(self, preconditioner: "Preconditioner"):
        self.preconditioner = preconditioner

    def __matmul__(self, x: torch.Tensor) -> torch.Tensor:
        return self.preconditioner._inverse_matmul(x)


class Preconditioner:
    def __init__(self, config: PreconditionerConfig):
        self.config = config

    # ---- to be provided by subclasses ----
    def _update(self, A, device: torch.device, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _matmul(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def _solve(self, x2d: torch.Tensor) -> torch.Tensor:  # pragma: no cover - abstract
        """P^{-1} applied to an (n, k) block."""
        raise NotImplementedError

    # ---- shared plumbing ----
    def _inverse_matmul_2d(self, x: torch.Tensor) -> torch.Tensor:
        return self._solve(x)

    def _inverse_matmul_1d(self, x: torch.Tensor) -> torch.Tensor:
        return self._solve(x.unsqueeze(-1)).squeeze(-1)

    def __matmul__(self, x: torch.Tensor) -> torch.Tensor:
        _is_torch_tensor_1d_2d(x, "x")
        return self._matmul(x)

    def _inverse_matmul(self, x: torch.Tensor) -> torch.Tensor:
        _is_torch_tensor_1d_2d(x, "x")
        return self._inverse_matmul_1d(x) if x.ndim == 1 else self._inverse_matmul_2d(x)

    def _inverse_matmul_compose(self, fn: Callable) -> Callable:
        return lambda *args, **kwargs: self._inverse_matmul(fn(*args, **kwargs))

    def _update_damping(self, baseline_rho: float):
        """No-op except for Nystrom with adaptive damping."""

    @property
    def _inv(self) -> _InvPreconditioner:
        return _InvPreconditioner(self)


class Identity(Preconditioner):
    """P = I (``identity.py:9-74``)."""

    def _update(self, A, device):
        pass

    def _matmul(self, x):
        return x

    def _solve(self, x2d):
        return x2d


class Newton(Preconditioner):
    """P = A + rho I through its Cholesky factor (``newton.py:8-88``).

    A linear operator is densified with one matmat against the identity (``newton.py:63``); unlike
    the reference, a dense ``A`` passed by the caller is not modified in place.
    """

    def __init__(self, config: NewtonConfig):
        super().__init__(config)
        self.L = None

    def _update(self, A, device):
        if isinstance(A, torch.Tensor):
            M = A.to(device).clone()
        else:
            M = A @ torch.eye(A.shape[1], dtype=A.dtype, device=device)
        M.diagonal().add_(self.config.rho)
        self.L = torch.linalg.cholesky(M, upper=False)

    def _matmul(self, x):
        return self.L @ (self.L.T @ x)

    def _solve(self, x2d):
        return torch.cholesky_solve(x2d, self.L, upper=False)


def _left_singular_vectors_gram(F: torch.Tensor, rows_per_chunk: int = 1 << 18):
    """Left singular vectors and squared singular values of a tall fp32 factor through its fp64 Gram matrix.

    ``G = F^T F`` is accumulated in fp64 (row chunks bound the temporary), ``eigh(G) = V diag(s^2) V^T`` and
    ``U = F V diag(1/s)`` is formed in fp64 and rounded once to fp32.  With fp64 carrying the squared condition
    number this resolves singular-value ratios down to ~1e-8 -- beyond what an fp32 QR + SVD resolves -- and costs
    one skinny GEMM pair plus an r x r ``syevd`` (C1: 14 ms -> 7 ms for the whole build; no tall ``geqrf``).
    """
    n, r = F.shape
    G = torch.zeros((r, r), dtype=torch.float64, device=F.device)
    for lo in range(0, n, rows_per_chunk):
        Fc = F[lo:lo + rows_per_chunk].double()
        G.addmm_(Fc.T, Fc)
    evals, V = torch.linalg.eigh(G)
    evals, V = evals.flip(0), V.flip(1)  # descending, like an SVD
    sig2 = evals.clamp_min(0.0)
    inv_sig = torch.where(sig2 > 0, sig2.clamp_min(torch.finfo(torch.float64).tiny).rsqrt(), torch.zeros_like(sig2))
    W = V * inv_sig  # (r, r)
    U = torch.empty_like(F)
    for lo in range(0, n, rows_per_chunk):
        U[lo:lo + rows_per_chunk] = (F[lo:lo + rows_per_chunk].double() @ W).to(F.dtype)
    return U, sig2.to(F.dtype)


class Nystrom(Preconditioner):
    """Randomized Nystrom approximation ``A ~= U diag(S) U^T``, P = U diag(S) U^T + rho I.

    Construction as in ``nystrom.py:55-98`` (sketch, eps-trace shift, Cholesky of the core,
    triangular solve, singular vectors, ``S = max(sigma^2 - shift, 0)``).  The singular vectors
    of the tall factor ``Y L^{-T}`` (n x r) are taken from its thin QR followed by the SVD of the
    r x r triangle: on the GPU that is one ``geqrf`` and an SVD independent of n, instead of a
    tall-skinny ``gesvd``.  Inverse: Woodbury in fp64, and the Cholesky-stabilised form of
    ``nystrom.py:113-127`` in lower precision.  For fp32 operators the singular vectors come from the fp64 Gram
    matrix of the factor instead (``_left_singular_vectors_gram``).
    """

    def __init__(self, config: NystromConfig):
        super().__init__(config)
        self.U = None
        self.S = None
        self.L = None
        self.low_precision = False

    def _update(self, A, device):
        self.low_precision = A.dtype != torch.float64
        self.L = None
        Omega = get_sketch(self.config.sketch, "right", self.config.rank, A.shape[1], dtype=A.dtype, device=device)
        Y = Omega._apply_right(A)  # (n, r): the kernel matmat with r right-hand sides
        core = Omega._apply_left_trans(Y)  # (r, r)
        shift = torch.finfo(Y.dtype).eps * torch.trace(core)
        core.diagonal().add_(shift)
        C = torch.linalg.cholesky(core, upper=False)
        # F = Y C^{-T}  (n, r), then A_nys = F F^T - shift I on range(F)
        F = torch.linalg.solve_triangular(C.T, Y, upper=True, left=False)
        if self.low_precision:
            self.U, sig2 = _left_singular_vectors_gram(F)
        else:
            Q, R = torch.linalg.qr(F, mode="reduced")
            Ur, sig, _ = torch.linalg.svd(R, full_matrices=False)
            self.U, sig2 = Q @ Ur, sig * sig
        self.S = torch.clamp(sig2 - shift, min=0.0)

    def _matmul(self, x):
        S = self.S if x.ndim == 1 else self.S.unsqueeze(-1)
        return self.U @ (S * (self.U.T @ x)) + self.config.rho * x

    def _solve(self, x2d):
        rho = self.config.rho
        UTx = self.U.T @ x2d
        if self.low_precision:
            if self.L is None:
                G = self.U.T @ self.U
                # S entries clamped to 0 would put inf on the diagonal (LinAlgError in the reference)
                G.diagonal().add_(rho / self.S.clamp_min(torch.finfo(self.S.dtype).tiny))
                self.L = torch.linalg.cholesky(G)
            return (x2d - self.U @ torch.cholesky_solve(UTx, self.L, upper=False)) / rho
        return (x2d - self.U @ UTx) / rho + self.U @ (UTx / (self.S + rho).unsqueeze(-1))

    def _update_damping(self, baseline_rho: float) -> None:
        if self.config.damping_mode == _DampingMode.ADAPTIVE:
            self.config.rho = baseline_rho + self.S[-1]
            self.L = None


_REGISTRY = {IdentityConfig: Identity, NewtonConfig: Newton, NystromConfig: Nystrom}


def _get_precond(precond_config: PreconditionerConfig) -> Preconditioner:
    """Instantiate the preconditioner that belongs to a config object (``factory.py:31-68``)."""
    cls = _REGISTRY.get(type(precond_config))
    if cls is None:
        if isinstance(precond_config, SkPreConfig):
            raise NotImplementedError("SkPre needs the reference's sparse sketches (out of scope for this package)")
        raise KeyError(f"No preconditioner found for configuration: {type(precond_config)}")
    return cls(precond_config)
