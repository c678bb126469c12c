"""Multi-device linear operators (interface of ``rlaopt/linops/distributed.py:15-208``).

ROW mode     matvec  = broadcast x, concatenate shard results   (``distributed.py:41-45``)
             rmatvec = scatter x by shard rows, sum             (``:82-86``)
COLUMN mode  matvec  = scatter x by shard columns, sum          (``:46-50``)
             rmatvec = broadcast x, concatenate                 (``:87-91``)

Single process, all devices driven asynchronously (see ``base.py``).
"""
from __future__ import annotations

import torch

from .base import _BaseLinOp, _BaseDistributedLinOp
from .enums import _DistributionMode, _Operation

__all__ = ["DistributedLinOp", "DistributedTwoSidedLinOp", "DistributedSymmetricLinOp"]


class _DistributedLinOp(_BaseDistributedLinOp):
    def _matvec(self, w: torch.Tensor) -> torch.Tensor:
        row_mode = self._distribution_mode == _DistributionMode.ROW
        # ROW: every shard sees all of w; COLUMN: shard i sees the rows of w matching its columns
        parts = self._run_shards(w, _Operation.MATVEC, chunk=not row_mode, by_dimension=1)
        return self._combine_results(parts, concatenate=row_mode, device=w.device)

    def _matmat(self, w: torch.Tensor) -> torch.Tensor:
        return self._matvec(w)


class _DistributedTwoSidedLinOp(_DistributedLinOp):
    def _rmatvec(self, w: torch.Tensor) -> torch.Tensor:
        row_mode = self._distribution_mode == _DistributionMode.ROW
        # ROW: shard i sees the rows of w matching its rows, partial results are summed
        parts = self._run_shards(w, _Operation.RMATVEC, chunk=row_mode, by_dimension=0)
        return self._combine_results(parts, concatenate=not row_mode, device=w.device)

    def _rmatmat(self, w: torch.Tensor) -> torch.Tensor:
        return self._rmatvec(w)

    @property
    def T(self) -> "_DistributedTwoSidedLinOp":
        return _DistributedTwoSidedLinOp(
            shape=torch.Size((self.shape[1], self.shape[0])),
            A=[op.T for op in self._A],
            distribution_mode=self._distribution_mode.flipped(),
            is_new=False,
        )


class _DistributedSymmetricLinOp(_DistributedTwoSidedLinOp):
    def __init__(self, shape, A, distribution_mode, is_new=True, **shared):
        super().__init__(shape=shape, A=A, distribution_mode=distribution_mode, is_new=is_new, **shared)
        if is_new and shape[0] != shape[1]:
            raise ValueError(
                "DistributedSymmetricLinOp requires the shape to be square. "
                f"The received shape is {shape}."
            )

    def _rmatvec(self, w: torch.Tensor) -> torch.Tensor:
        return self._matvec(w)

    def _rmatmat(self, w: torch.Tensor) -> torch.Tensor:
        return self._matmat(w)

    @property
    def T(self) -> "_DistributedSymmetricLinOp":
        return self


class DistributedLinOp(_DistributedLinOp):
    """Operator sharded over devices; ``A`` lists one shard operator per device."""

    def __init__(self, shape: torch.Size, A: list[_BaseLinOp], distribution_mode: str):
        super().__init__(shape=shape, A=A, distribution_mode=distribution_mode, is_new=True)


class DistributedTwoSidedLinOp(_DistributedTwoSidedLinOp):
    """Sharded operator whose shards also support ``.T``."""

    def __init__(self, shape: torch.Size, A: list[_BaseLinOp], distribution_mode: str):
        super().__init__(shape=shape, A=A, distribution_mode=distribution_mode, is_new=True)


class DistributedSymmetricLinOp(_DistributedSymmetricLinOp):
    """Sharded symmetric operator (square; ``.T`` is the operator itself)."""

    def __init__(self, shape: torch.Size, A: list[_BaseLinOp], distribution_mode: str):
        super().__init__(shape=shape, A=A, distribution_mode=distribution_mode, is_new=True)
