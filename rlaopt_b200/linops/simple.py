"""Single-device linear operators assembled from user callables.

Public interface of ``rlaopt/linops/simple.py:15-104``:

* ``LinOp(device, shape, matvec, matmat=None, dtype)``                       -- forward products only
* ``TwoSidedLinOp(device, shape, matvec, rmatvec, matmat=None, rmatmat=None, dtype)`` -- adds ``x @ A`` and ``.T``
* ``SymmetricLinOp(device, shape, matvec, matmat=None, dtype)``              -- square, ``A.T is A``

Internally one table of four product slots (``fwd_vec``, ``fwd_mat``, ``adj_vec``, ``adj_mat``) backs all three
classes; a block product that the caller did not supply is derived from the vector product by mapping over
the columns (``torch.vmap``, as the reference does at ``simple.py:32,62``).  Transposition swaps the forward and
adjoint slots and -- unlike the reference (``simple.py:73-81``) -- carries the dtype along.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import torch

from rlaopt_b200.utils import _is_callable

from .base import _BaseLinOp

__all__ = ["LinOp", "TwoSidedLinOp", "SymmetricLinOp"]

_DEFAULT_DTYPE = torch.get_default_dtype()  # frozen at import, like the reference's module-level default


class _Products(NamedTuple):
    fwd_vec: Callable
    fwd_mat: Callable
    adj_vec: Optional[Callable] = None
    adj_mat: Optional[Callable] = None

    def swapped(self) -> "_Products":
        return _Products(self.adj_vec, self.adj_mat, self.fwd_vec, self.fwd_mat)


def _slot_pair(vec: Callable, mat: Optional[Callable], vec_name: str, mat_name: str):
    """Validate one (vector product, block product) pair; derive the block product when it is missing."""
    _is_callable(vec, vec_name)
    if mat is None:
        return vec, torch.vmap(vec, in_dims=1, out_dims=1)
    _is_callable(mat, mat_name)
    return vec, mat


class LinOp(_BaseLinOp):
    def __init__(self, device: torch.device, shape: torch.Size, matvec: Callable, matmat: Optional[Callable] = None,
                 dtype: torch.dtype = _DEFAULT_DTYPE):
        super().__init__(device=device, shape=shape, dtype=dtype)
        self._ops = _Products(*_slot_pair(matvec, matmat, "matvec", "matmat"))

    def _matvec(self, x: torch.Tensor) -> torch.Tensor:
        return self._ops.fwd_vec(x)

    def _matmat(self, x: torch.Tensor) -> torch.Tensor:
        return self._ops.fwd_mat(x)


class TwoSidedLinOp(LinOp):
    def __init__(self, device: torch.device, shape: torch.Size, matvec: Callable, rmatvec: Callable,
                 matmat: Optional[Callable] = None, rmatmat: Optional[Callable] = None,
                 dtype: torch.dtype = _DEFAULT_DTYPE):
        super().__init__(device, shape, matvec, matmat, dtype)
        self._ops = self._ops._replace(**dict(zip(("adj_vec", "adj_mat"),
                                                   _slot_pair(rmatvec, rmatmat, "rmatvec", "rmatmat"))))

    def _rmatvec(self, x: torch.Tensor) -> torch.Tensor:
        return self._ops.adj_vec(x)

    def _rmatmat(self, x: torch.Tensor) -> torch.Tensor:
        return self._ops.adj_mat(x)

    @property
    def T(self) -> "TwoSidedLinOp":
        t = self._ops.swapped()
        rows, cols = self.shape
        return TwoSidedLinOp(self.device, torch.Size((cols, rows)), t.fwd_vec, t.adj_vec, t.fwd_mat, t.adj_mat,
                             self.dtype)


class SymmetricLinOp(TwoSidedLinOp):
    def __init__(self, device: torch.device, shape: torch.Size, matvec: Callable, matmat: Optional[Callable] = None,
                 dtype: torch.dtype = _DEFAULT_DTYPE):
        super().__init__(device, shape, matvec, matvec, matmat, matmat, dtype)
        if shape[0] != shape[1]:
            raise ValueError(f"SymmetricLinOp requires the shape to be square. The received shape is {shape}.")

    @property
    def T(self) -> "SymmetricLinOp":
        return self
