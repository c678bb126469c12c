"""Fused products: ``A @ x`` with the caller's next passes folded into the operator's output stage.

    Y    = alpha * (A @ x) + beta * addend[addend_idx] + gamma * rhs[rhs_idx]
    gram = gram_with^T Y            (optional)
    sqn  = column sums of Y^2       (optional)

Kernel operators implement ``matmat_fused`` on the GPU (``rlaopt_b200.ops.matmat_packed_fused``: the terms are added
where the column-split partial sums are reduced, the Gram matrix and the norms are accumulated in that same pass).
Any other operator -- a dense tensor, a user ``LinOp``, the reference-compatible multi-device classes -- gets the
same result from separate passes, so the solvers are written once against :func:`apply_fused`:

* ``A P + reg P`` and ``P^T A P``                     ``rlaopt/solvers/pcg.py:58-61``
* ``B - (A W + reg W)`` and its column norms          ``rlaopt/models/linsys.py:96-99``, ``solvers/pcg.py:33``
* ``A[blk, :] Y + reg Y[blk] - B[blk]``               ``rlaopt/solvers/sap.py:113-127``
"""
from __future__ import annotations

from typing import Optional

import torch

__all__ = ["apply_fused"]


def _take(t: torch.Tensor, idx: Optional[torch.Tensor]) -> torch.Tensor:
    return t if idx is None else t[idx.to(t.device)]


def apply_fused(
    A,
    x: torch.Tensor,
    *,
    alpha: float = 1.0,
    addend: Optional[torch.Tensor] = None,
    beta: float = 0.0,
    addend_idx: Optional[torch.Tensor] = None,
    rhs: Optional[torch.Tensor] = None,
    gamma: float = 0.0,
    rhs_idx: Optional[torch.Tensor] = None,
    gram_with: Optional[torch.Tensor] = None,
    want_sqnorm: bool = False,
    store: bool = True,
):
    """Returns ``(Y or None, gram or None, sqnorm or None)``; ``store=False`` asks for the reductions only."""
    k = 1 if x.ndim == 1 else x.shape[1]
    g = 0 if gram_with is None else (1 if gram_with.ndim == 1 else gram_with.shape[1])
    fused = getattr(A, "matmat_fused", None)
    if fused is not None and ((gram_with is None and not want_sqnorm) or getattr(A, "fused_reductions_ok", lambda *_: False)(k, g)):
        return fused(x, alpha=alpha, addend=addend, beta=beta, addend_idx=addend_idx, rhs=rhs, gamma=gamma,
                     rhs_idx=rhs_idx, gram_with=gram_with, want_sqnorm=want_sqnorm, store=store)
    if fused is not None:  # reductions beyond the fused stage's range: fuse the element-wise terms, reduce separately
        Y, _, _ = fused(x, alpha=alpha, addend=addend, beta=beta, addend_idx=addend_idx, rhs=rhs, gamma=gamma,
                        rhs_idx=rhs_idx)
    else:
        Y = A @ x
        if alpha != 1.0:
            Y = Y * alpha
        if addend is not None:
            Y = Y.add(_take(addend, addend_idx).reshape(Y.shape), alpha=beta)
        if rhs is not None:
            Y = Y.add(_take(rhs, rhs_idx).reshape(Y.shape), alpha=gamma)
    Y2 = Y.unsqueeze(1) if Y.ndim == 1 else Y
    gram = None
    if gram_with is not None:
        L = gram_with.unsqueeze(1) if gram_with.ndim == 1 else gram_with
        gram = L.T @ Y2
    sqn = (Y2 * Y2).sum(dim=0) if want_sqnorm else None
    return (Y if store else None), gram, sqn
