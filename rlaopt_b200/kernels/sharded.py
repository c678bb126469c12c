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

from rlaopt_b200.linops.spmd import RowShardedLinOp, shard_rows

from .base import _KernelLinOp
from .configs import KernelConfig

__all__ = ["sharded_kernel_linop", "replicate_from_host"]


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
) -> RowShardedLinOp:
    """Row-sharded ``c * K(A1, A2)`` for kernel name ``kernel`` ("rbf", "laplace", "matern12|32|52")."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, m = A1.shape[0], A2.shape[0]
    lo, hi = shard_rows(n, world)[rank]
    same = A1 is A2 or (A1.data_ptr() == A2.data_ptr() and A1.shape == A2.shape and A1.device == A2.device)
    A2_dev = A2.to(device)
    local = None
    if hi > lo:
        A1_dev = A2_dev[lo:hi] if same else A1[lo:hi].to(device)
        local = _KernelLinOp(A1_dev, A2_dev, kernel_config.to(device), _kernel_key=kernel.lower())
    return RowShardedLinOp(local, torch.Size((n, m)), device, A1.dtype, group)
