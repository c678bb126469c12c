"""Gaussian and orthonormal sketching matrices.

``Omega_mat`` has shape ``(sketch_size, matrix_dim)`` in ``"left"`` mode and
``(matrix_dim, sketch_size)`` in ``"right"`` mode.  The random draws have the shapes and order of
the reference (``gauss.py:46-52``: ``randn(s, d) / sqrt(s)``; ``ortho.py:50-56``: reduced QR of
``randn(d, s)``) so that a seeded run under ``rlaopt_b200.utils.host_rng`` consumes the same
random stream; on the device the QR runs through cuSOLVER.
"""
from __future__ import annotations

import torch

from rlaopt_b200.utils import _is_pos_int, randn

_SIDES = ("left", "right")
_KINDS = ("gauss", "ortho", "sparse")


def _choice(value, options, param_name):
    if isinstance(value, str) and value.lower() in options:
        return value.lower()
    raise ValueError(f"Invalid value for {param_name}: {value}. Expected one of {list(options)}.")


class Sketch:
    """Base class: holds ``Omega_mat`` and applies it from either side."""

    def __init__(self, mode: str, sketch_size: int, matrix_dim: int, dtype: torch.dtype, device: torch.device):
        self.mode = _choice(mode, _SIDES, "mode")
        _is_pos_int(sketch_size, "sketch_size")
        self.s, self.d = sketch_size, matrix_dim
        self.dtype, self.device = dtype, device
        self.Omega_mat = self._generate_embedding()

    def _generate_embedding(self) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    # x may be a tensor or a LinOp: ``LinOp @ tensor`` and ``tensor @ LinOp`` both dispatch to the fused matmat
    def _apply_left(self, x):
        return self.Omega_mat @ x

    def _apply_right(self, x):
        return x @ self.Omega_mat

    def _apply_left_trans(self, x):
        return self.Omega_mat.T @ x

    def _apply_right_trans(self, x):
        return x @ self.Omega_mat.T


class Gauss(Sketch):
    """i.i.d. N(0, 1/s) entries, so that ``Omega.T @ Omega`` is an isometry in expectation."""

    def _generate_embedding(self) -> torch.Tensor:
        G = randn(self.s, self.d, dtype=self.dtype, device=self.device) / self.s**0.5
        return (G.T if self.mode == "right" else G).contiguous()


class Ortho(Sketch):
    """Orthonormal columns (right mode) / rows (left mode): Q factor of a Gaussian matrix."""

    def _generate_embedding(self) -> torch.Tensor:
        Q = torch.linalg.qr(randn(self.d, self.s, dtype=self.dtype, device=self.device), mode="reduced")[0]
        return (Q.T if self.mode == "left" else Q).contiguous()


def get_sketch(name: str, mode: str, sketch_size: int, matrix_dim: int, dtype: torch.dtype,
               device: torch.device) -> Sketch:
    """Factory with the reference's signature (``sketches/factory.py:26-59``)."""
    kind = _choice(name, _KINDS, "name")
    if kind == "sparse":
        raise NotImplementedError(
            "the sparse sign sketch multiplies through the reference's CSC kernels, which are outside the "
            "kernel-matmat path rebuilt here; use 'gauss' or 'ortho'")
    return (Gauss if kind == "gauss" else Ortho)(mode, sketch_size, matrix_dim, dtype, device)
