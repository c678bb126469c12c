"""SPMD (one process per GPU) kernel operator: rows of A1 sharded over the ranks.

Under ``torchrun`` every rank calls :func:`sharded_kernel_linop` with the same
point sets; the rank keeps its row block of ``A1`` and a replica of ``A2`` on its
own GPU (``rlaopt/kernels/base.py:143-144,297-307``), applies the fused CUDA kernel to
its block, and the row blocks are all-gathered over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from rlaopt_b200 import ops
from rlaopt_b200.linops import LinOp
from rlaopt_b200.linops.spmd import RowShardedLinOp, shard_rows

from .base import _KernelLinOp
from .configs import KernelConfig

__all__ = ["sharded_kernel_linop", "replicate_from_host", "ShardedKernelLinOp"]


class ShardedKernelLinOp(RowShardedLinOp):
    """Row-sharded kernel operator with the oracles SAP / ASkotch needs, distributed as the reference does it
    (``rlaopt/kernels/base.py:408-505``):

    * ``row_oracle(blk)``  -- COLUMN mode: every rank owns a chunk of the columns ``A2[lo:hi]``, multiplies
      ``c K(A1[blk], A2[lo:hi])`` with its rows of ``x`` and the partial results are summed (all-reduce of b x k);
    * ``blk_oracle(blk)``  -- ROW mode over ``torch.chunk(blk, world)``: every rank computes its rows of
      ``c K(A1[blk], A2[blk]) @ x`` and the row blocks are all-gathered.

    ``A1`` and ``A2`` are replicated on every rank (one copy per GPU, as ``base.py:143-144``); inputs and outputs of
    all products are replicated tensors.
    """

    def __init__(self, A1_dev, A2_dev, kernel_config, kernel, device, group=None):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n, m = A1_dev.shape[0], A2_dev.shape[0]
        lo, hi = shard_rows(n, world)[rank]
        self._kernel = kernel.lower()
        self._cfg = kernel_config.to(device)
        self._A1, self._A2 = A1_dev, A2_dev
        local = None
        if hi > lo:
            local = _KernelLinOp(A1_dev[lo:hi], A2_dev, self._cfg, _kernel_key=self._kernel)
        super().__init__(local, torch.Size((n, m)), device, A1_dev.dtype, group)
        # column chunk of A2 owned by this rank (row oracle) -- its pack is built once, on first use
        self._clo, self._chi = shard_rows(m, world)[rank]
        self._col_op = None
        if self._chi > self._clo:
            self._col_op = _KernelLinOp(A1_dev, A2_dev[self._clo:self._chi], self._cfg, _kernel_key=self._kernel)

    A1 = property(lambda self: self._A1)
    A2 = property(lambda self: self._A2)
    kernel_config = property(lambda self: self._cfg)

    def row_oracle(self, blk: torch.Tensor) -> LinOp:
        blk_dev = blk.to(self.device)
        local = self._col_op.row_oracle(blk_dev) if self._col_op is not None else None
        clo, chi, group = self._clo, self._chi, self.group

        def matmat(x: torch.Tensor) -> torch.Tensor:
            if local is None:
                part = x.new_zeros((blk_dev.shape[0],) + tuple(x.shape[1:]))
            else:
                part = (local @ x[clo:chi]).contiguous()
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
            return part

        def matmat_fused(x, *, alpha=1.0, addend=None, beta=0.0, addend_idx=None, rhs=None, gamma=0.0, rhs_idx=None,
                         gram_with=None, want_sqnorm=False, store=True):
            """Block gradient of ASkotch (``sap.py:113-127``): the partial products are summed over the ranks, so the
            ``beta`` / ``gamma`` terms are added by rank 0's output stage only and ride the same all-reduce."""
            if gram_with is not None or want_sqnorm or not store:
                raise NotImplementedError("the distributed row oracle fuses the element-wise terms only")
            first = self.rank == 0
            if local is None:
                part = x.new_zeros((blk_dev.shape[0],) + tuple(x.shape[1:]))
                if first:
                    for t, idx, coef in ((addend, addend_idx, beta), (rhs, rhs_idx, gamma)):
                        if t is not None:
                            rows_t = t if idx is None else t[idx.to(self.device)]
                            part.add_(rows_t.reshape(part.shape), alpha=coef)
            else:
                part, _, _ = local.matmat_fused(
                    x[clo:chi], alpha=alpha, addend=addend if first else None, beta=beta, addend_idx=addend_idx,
                    rhs=rhs if first else None, gamma=gamma, rhs_idx=rhs_idx)
                part = part.contiguous()
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
            return part, None, None

        op = LinOp(self.device, torch.Size((blk_dev.shape[0], self.shape[1])), matmat, matmat, dtype=self.dtype)
        op.matmat_fused = matmat_fused
        op.fused_reductions_ok = lambda k, g=0: False
        return op

    def blk_oracle(self, blk: torch.Tensor) -> LinOp:
        blk_dev = blk.to(self.device)
        b = blk_dev.shape[0]
        lo, hi = shard_rows(b, self.world)[self.rank]
        mine = blk_dev[lo:hi]
        block = -(-b // self.world)
        cfg, A1, A2, group, world = self._cfg, self._A1, self._A2, self.group, self.world
        kid = ops.kernel_id(self._kernel)
        packs: dict[int, tuple] = {}  # per layout: this rank's rows of the block and the whole block, packed once

        def block_packs(k: int):
            layout = ops.choose_layout(kid, A1.dtype, A1.shape[1], k)
            if layout not in packs:
                def build(lay):
                    c = ops.column_mean(A2, blk_dev) if lay == ops.LAYOUT_TC else None
                    return (ops.pack_points(A1, cfg.lengthscale, mine, lay, c),
                            ops.pack_points(A2, cfg.lengthscale, blk_dev, lay, c))
                Pr, Pc = build(layout)
                if layout == ops.LAYOUT_TC and not ops.tc_accuracy_ok(kid, Pr.max_sqnorm, Pc.max_sqnorm):
                    Pr, Pc = build(ops.LAYOUT_SIMT)
                packs[layout] = (Pr, Pc)
            return packs[layout]

        def matmat(x: torch.Tensor) -> torch.Tensor:
            vec = x.ndim == 1
            xm = x.unsqueeze(1) if vec else x
            buf = xm.new_zeros((world * block, xm.shape[1]))
            if hi > lo:
                Pr, Pc = block_packs(xm.shape[1])
                buf[self.rank * block:self.rank * block + (hi - lo)] = ops.matmat_packed(
                    Pr, Pc, xm, kid, cfg.const_scaling)
            dist.all_gather_into_tensor(buf, buf[self.rank * block:(self.rank + 1) * block], group=group)
            if world * block != b:  # ragged: drop the padding rows of every rank's slot
                ranges = shard_rows(b, world)
                buf = torch.cat([buf[r * block:r * block + (h - l)] for r, (l, h) in enumerate(ranges)])
            return buf[:, 0] if vec else buf

        return LinOp(self.device, torch.Size((b, b)), matmat, matmat, dtype=self.dtype)


def replicate_from_host(T_host: torch.Tensor, device: torch.device,
                        group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Replica of a (pinned) host matrix on every rank's GPU: each rank copies 1/world of the rows over
    PCIe and the blocks are all-gathered over NVLink, instead of every rank pulling the whole matrix
    through the host link (``rlaopt/kernels/base.py:143-144`` moves the full A2 to every device)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return T_host.to(device, non_blocking=True)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = T_host.shape[0]
    block = -(-n // world)
    full = torch.empty((block * world,) + tuple(T_host.shape[1:]), dtype=T_host.dtype, device=device)
    lo, hi = min(rank * block, n), min((rank + 1) * block, n)
    mine = full[rank * block:rank * block + (hi - lo)]
    mine.copy_(T_host[lo:hi], non_blocking=True)
    # in-place all-gather: this rank's block already sits at its slot of the output (ragged tails are padding)
    dist.all_gather_into_tensor(full, full[rank * block:(rank + 1) * block], group=group)
    return full[:n]


def sharded_kernel_linop(
    A1: torch.Tensor,
    A2: torch.Tensor,
    kernel_config: KernelConfig,
    kernel: str,
    device: torch.device,
    group: Optional[dist.ProcessGroup] = None,
) -> ShardedKernelLinOp:
    """Row-sharded ``c * K(A1, A2)`` (with SPMD row / block oracles) for kernel name ``kernel`` ("rbf", "laplace", "matern12|32|52")."""
    same = A1 is A2 or (A1.data_ptr() == A2.data_ptr() and A1.shape == A2.shape and A1.device == A2.device)
    A2_dev = A2.to(device)
    A1_dev = A2_dev if same else A1.to(device)
    return ShardedKernelLinOp(A1_dev, A2_dev, kernel_config, kernel, device, group)
