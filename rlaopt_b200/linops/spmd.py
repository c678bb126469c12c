"""One-process-per-GPU (SPMD) row-sharded operator over ``torch.distributed``.

The reference's multi-device layer is host-side multiprocessing
(``rlaopt/linops/base.py:114-291``); its only collective prototype is
``experiments/distributed_matvec_v4.py:36-79`` (NCCL ``all_gather`` of per-rank
partial results).  This is the production form of that prototype for launches
under ``torchrun``: every rank owns the row block ``torch.chunk(arange(n), world)[rank]``
of the operator (same partition as ``rlaopt/kernels/base.py:297-299``),

    matvec  : local block product, then all-gather of the row blocks  (ROW mode, concat)
    rmatvec : local partial product of the rank's rows, then all-reduce (ROW mode, sum)

Backend-agnostic: NCCL over NVLink on GPUs, gloo on CPU for the tests.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .base import _BaseLinOp

__all__ = ["RowShardedLinOp", "shard_rows"]


def shard_rows(n: int, world: int) -> list[tuple[int, int]]:
    """[lo, hi) row ranges per rank — ``torch.chunk(arange(n), world)`` without the tensor.

    ``torch.chunk`` may produce fewer than ``world`` chunks; the remaining ranks own
    an empty range.
    """
    size = -(-n // world)  # ceil
    ranges = []
    for r in range(world):
        lo = min(r * size, n)
        ranges.append((lo, min(lo + size, n)))
    return ranges


class RowShardedLinOp(_BaseLinOp):
    """Global (n x m) operator whose rows are sharded over the ranks of a process group.

    ``local_op`` is this rank's (n_r x m) block (``None`` for a rank that owns no
    rows).  Inputs and outputs are replicated: every rank passes the same ``x`` and
    receives the full result, which keeps the solvers' code unchanged.
    """

    def __init__(
        self,
        local_op: Optional[_BaseLinOp],
        shape: torch.Size,
        device: torch.device,
        dtype: torch.dtype,
        group: Optional[dist.ProcessGroup] = None,
    ):
        super().__init__(device=device, shape=shape, dtype=dtype)
        if not dist.is_initialized():
            raise RuntimeError("RowShardedLinOp needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ranges = shard_rows(shape[0], self.world)
        self.lo, self.hi = self.ranges[self.rank]
        if local_op is None and self.hi > self.lo:
            raise ValueError("this rank owns rows but has no local operator")
        if local_op is not None and tuple(local_op.shape) != (self.hi - self.lo, shape[1]):
            raise ValueError(f"local block has shape {tuple(local_op.shape)}, expected {(self.hi - self.lo, shape[1])}")
        self.local_op = local_op
        self._block = self.ranges[0][1] - self.ranges[0][0]  # largest block (padding size)

    # -- local pieces (no communication) -----------------------------------
    def local_matmat(self, x: torch.Tensor) -> torch.Tensor:
        """This rank's rows of ``A @ x``."""
        cols = 1 if x.ndim == 1 else x.shape[1]
        if self.local_op is None:
            return x.new_zeros((0,) if x.ndim == 1 else (0, cols))
        return self.local_op @ x

    # -- global products ---------------------------------------------------
    def _gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        vec = local.ndim == 1
        loc = local.unsqueeze(1) if vec else local
        k = loc.shape[1]
        if loc.shape[0] < self._block:
            pad = loc.new_zeros((self._block - loc.shape[0], k))
            loc = torch.cat([loc, pad], dim=0)
        buf = loc.new_empty((self.world * self._block, k))
        dist.all_gather_into_tensor(buf, loc.contiguous(), group=self.group)
        n = self.shape[0]
        if self.world * self._block != n:  # ragged tail: drop the padding rows
            buf = torch.cat([buf[r * self._block : r * self._block + (hi - lo)] for r, (lo, hi) in enumerate(self.ranges)])
        return buf[:, 0] if vec else buf

    def _matvec(self, x: torch.Tensor) -> torch.Tensor:
        return self._gather_rows(self.local_matmat(x))

    def matmat_to_host(self, x: torch.Tensor, host_out: torch.Tensor, wait: bool = True) -> torch.Tensor:
        """``A @ x`` delivered to host memory that every rank maps (``rlaopt_b200.utils.SharedPinnedTensor``): each
        rank copies its own row block over its own PCIe link -- no gather on one GPU, no single-link funnel.  With
        ``wait`` the call returns once all ranks' blocks have landed."""
        vec = x.ndim == 1
        loc = self.local_matmat(x)
        if self.hi > self.lo:
            dst = host_out[self.lo:self.hi]
            dst.copy_(loc.reshape(dst.shape), non_blocking=True)
        if wait:
            if self._device.type == "cuda":
                torch.cuda.current_stream(self._device).synchronize()
            dist.barrier(group=self.group)
        return host_out[:, 0] if (vec and host_out.ndim == 2) else host_out

    # -- fused products (rlaopt_b200.linops.apply_fused) ---------------------
    def fused_reductions_ok(self, k: int, gram_cols: int = 0) -> bool:
        # must not depend on the rank (a rank without rows would otherwise take another collective path)
        return True

    def matmat_fused(self, x: torch.Tensor, *, alpha=1.0, addend=None, beta=0.0, addend_idx=None, rhs=None, gamma=0.0,
                     rhs_idx=None, gram_with=None, want_sqnorm=False, store=True):
        """Every rank runs the fused output stage on its own row block (its rows of ``addend`` / ``rhs`` /
        ``gram_with``); the row blocks are all-gathered and the k x k Gram partials and the column norms are
        all-reduced together in one small buffer -- the only reductions block PCG needs across ranks
        (``rlaopt/solvers/pcg.py:58-61``)."""
        from .fused import apply_fused

        lo, hi = self.lo, self.hi
        k = 1 if x.ndim == 1 else x.shape[1]
        g = 0 if gram_with is None else (1 if gram_with.ndim == 1 else gram_with.shape[1])

        def rows_of(t, idx):
            if t is None:
                return None, None
            if idx is None:
                return t[lo:hi], None
            return t, idx.to(t.device)[lo:hi]

        Y_loc = gram = sqn = None
        if self.local_op is not None:
            a_t, a_i = rows_of(addend, addend_idx)
            r_t, r_i = rows_of(rhs, rhs_idx)
            # the local operator's own fused stage when it has one (kernel operators), separate passes otherwise
            Y_loc, gram, sqn = apply_fused(
                self.local_op, x, alpha=alpha, addend=a_t, beta=beta, addend_idx=a_i, rhs=r_t, gamma=gamma, rhs_idx=r_i,
                gram_with=None if gram_with is None else gram_with[lo:hi], want_sqnorm=want_sqnorm, store=store)
        if g or want_sqnorm:
            red = x.new_zeros((g + (1 if want_sqnorm else 0)) * k)
            if gram is not None:
                red[: g * k] = gram.reshape(-1)
            if sqn is not None:
                red[g * k:] = sqn
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.group)
            gram = red[: g * k].reshape(g, k) if g else None
            sqn = red[g * k:] if want_sqnorm else None
        Y = None
        if store:
            if Y_loc is None:
                Y_loc = x.new_zeros((0,) if x.ndim == 1 else (0, k))
            Y = self._gather_rows(Y_loc)
        return Y, gram, sqn

    def _matmat(self, x: torch.Tensor) -> torch.Tensor:
        return self._matvec(x)

    def _rmatvec(self, w: torch.Tensor) -> torch.Tensor:
        m = self.shape[1]
        if self.local_op is None:
            part = w.new_zeros((m,) if w.ndim == 1 else (m, w.shape[1]))
        else:
            part = (self.local_op.T @ w[self.lo : self.hi]).contiguous()
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return part

    def _rmatmat(self, w: torch.Tensor) -> torch.Tensor:
        return self._rmatvec(w)

    @property
    def T(self):
        outer = self

        class _Adjoint(_BaseLinOp):
            def _matvec(self, x):
                return outer._rmatvec(x)

            def _matmat(self, x):
                return outer._rmatvec(x)

            def _rmatvec(self, x):
                return outer._matvec(x)

            def _rmatmat(self, x):
                return outer._matvec(x)

            @property
            def T(self):
                return outer

        return _Adjoint(self._device, torch.Size((self.shape[1], self.shape[0])), self._dtype)
