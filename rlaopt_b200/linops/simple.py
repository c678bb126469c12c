"""Single-device linear operators built from callables.

Interface of ``rlaopt/linops/simple.py:15-104``: ``LinOp(device, shape, matvec,
matmat=None, dtype)``, ``TwoSidedLinOp(..., rmatvec, ..., rmatmat=None)`` with
``.T``, and ``SymmetricLinOp``.  A missing ``matmat`` is derived from ``matvec``
with ``torch.vmap`` over columns (``simple.py:32,62``).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .base import _BaseLinOp
from rlaopt_b200.utils import _is_callable

__all__ = ["LinOp", "TwoSidedLinOp", "SymmetricLinOp"]

_DEFAULT_DTYPE = torch.get_default_dtype()


def _columnwise(fn: Callable) -> Callable:
    return torch.vmap(fn, in_dims=1, out_dims=1)


class LinOp(_BaseLinOp):
    def __init__(
        self,
        device: torch.device,
        shape: torch.Size,
        matvec: Callable,
        matmat: Optional[Callable] = None,
        dtype: torch.dtype = _DEFAULT_DTYPE,
    ):
        super().__init__(device=device, shape=shape, dtype=dtype)
        _is_callable(matvec, "matvec")
        if matmat is not None:
            _is_callable(matmat, "matmat")
        self._matvec_fn = matvec
        self._matmat_fn = matmat if matmat is not None else _columnwise(matvec)

    def _matvec(self, x: torch.Tensor) -> torch.Tensor:
        return self._matvec_fn(x)

    def _matmat(self, x: torch.Tensor) -> torch.Tensor:
        return self._matmat_fn(x)


class TwoSidedLinOp(LinOp):
    def __init__(
        self,
        device: torch.device,
        shape: torch.Size,
        matvec: Callable,
        rmatvec: Callable,
        matmat: Optional[Callable] = None,
        rmatmat: Optional[Callable] = None,
        dtype: torch.dtype = _DEFAULT_DTYPE,
    ):
        super().__init__(device, shape, matvec, matmat, dtype)
        _is_callable(rmatvec, "rmatvec")
        if rmatmat is not None:
            _is_callable(rmatmat, "rmatmat")
        self._rmatvec_fn = rmatvec
        self._rmatmat_fn = rmatmat if rmatmat is not None else _columnwise(rmatvec)

    def _rmatvec(self, x: torch.Tensor) -> torch.Tensor:
        return self._rmatvec_fn(x)

    def _rmatmat(self, x: torch.Tensor) -> torch.Tensor:
        return self._rmatmat_fn(x)

    @property
    def T(self) -> "TwoSidedLinOp":
        # the reference drops the dtype here (simple.py:73-81); we keep it
        return TwoSidedLinOp(
            device=self.device,
            shape=torch.Size((self.shape[1], self.shape[0])),
            matvec=self._rmatvec,
            rmatvec=self._matvec,
            matmat=self._rmatmat,
            rmatmat=self._matmat,
            dtype=self.dtype,
        )


class SymmetricLinOp(TwoSidedLinOp):
    def __init__(
        self,
        device: torch.device,
        shape: torch.Size,
        matvec: Callable,
        matmat: Optional[Callable] = None,
        dtype: torch.dtype = _DEFAULT_DTYPE,
    ):
        super().__init__(device, shape, matvec, matvec, matmat, matmat, dtype)
        if shape[0] != shape[1]:
            raise ValueError(
                f"SymmetricLinOp requires the shape to be square. The received shape is {shape}."
            )

    @property
    def T(self) -> "SymmetricLinOp":
        return self
