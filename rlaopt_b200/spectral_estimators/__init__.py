"""Spectral estimators used by the SAP / ASkotch step size (mirror of ``rlaopt.spectral_estimators``)."""
from __future__ import annotations

import torch

from rlaopt_b200.utils import randn

__all__ = ["randomized_powering"]


def randomized_powering(A, max_iters: int = 10, rtol: float = 1e-3):
    """Largest eigenvalue of a symmetric operator by power iteration from a Gaussian start.

    Same iteration and stopping rule as ``rlaopt/spectral_estimators/spectral_norm.py:11-29``
    (stop once the Rayleigh quotient moves by less than ``rtol`` of its previous value); returns
    ``(eigenvalue estimate, unit vector)``.  The start vector is drawn in ``A.dtype`` (the
    reference draws it in the default dtype, which breaks fp64 operators -- SURVEY appendix A).
    """
    n = A.shape[0]
    v = randn(n, dtype=getattr(A, "dtype", None), device=A.device)
    v = v / torch.linalg.norm(v, 2)
    sig, sig_new, err, it = 0.0, None, float("inf"), 0
    while it < max_iters and err > rtol * sig:
        w = A @ v
        sig_new = torch.dot(v, w)
        v = w / torch.linalg.norm(w, 2)
        err, sig = torch.abs(sig_new - sig), sig_new
        it += 1
    return sig_new, v
