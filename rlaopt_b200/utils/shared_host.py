"""Page-locked host memory shared by the ranks of one node.

The reference's distributed operator returns its result to ONE calling process through host memory
(``rlaopt/linops/base.py:259-276``: every worker's block travels through a pickled CPU tensor).  Under SPMD the
natural equivalent is a host buffer that every rank maps and pins: each rank copies its own row block of the result
over its own PCIe link (``RowShardedLinOp.matmat_to_host``), instead of gathering the blocks on one GPU and pushing
all of them through that GPU's link.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["SharedPinnedTensor", "shared_host_available"]


def shared_host_available(nbytes: int, limit_bytes: int = 4 << 30) -> bool:
    """Whether ``/dev/shm`` can back ``nbytes`` of shared, page-locked memory (half of its free space at most, and
    not more than ``limit_bytes``: registration of very large shared mappings fails on some hosts)."""
    try:
        st = os.statvfs("/dev/shm")
    except OSError:
        return False
    return nbytes <= limit_bytes and nbytes <= (st.f_bavail * st.f_frsize) // 2


class SharedPinnedTensor:
    """A ``/dev/shm``-backed tensor mapped by every rank of ``group`` and registered with CUDA (pinned) in each.

    ``.tensor`` is an ordinary CPU tensor (same bytes in all ranks); ``close()`` unpins, and the creating rank removes
    the backing file.  Collective: every rank of the group must construct it with the same arguments.
    """

    def __init__(self, name: str, shape: Sequence[int], dtype: torch.dtype = torch.float32,
                 group: Optional[dist.ProcessGroup] = None):
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if multi else 0
        self.group = group if multi else None
        self._multi = multi
        numel = 1
        for s in shape:
            numel *= int(s)
        tag = os.environ.get("MASTER_PORT", str(os.getpid() if not multi else 0))
        self.path = os.path.join("/dev/shm", f"rlaopt_b200_{tag}_{name}")
        if self.rank == 0:
            if os.path.exists(self.path):
                os.remove(self.path)
            flat = torch.from_file(self.path, shared=True, size=numel, dtype=dtype)
            if multi:
                dist.barrier(group=group)
        else:
            dist.barrier(group=group)  # the file exists with its full size once rank 0 has mapped it
            flat = torch.from_file(self.path, shared=True, size=numel, dtype=dtype)
        self._flat = flat
        self.tensor = flat.view(*shape)
        self._pinned = False
        if torch.cuda.is_available() and numel > 0:
            rc = torch.cuda.cudart().cudaHostRegister(flat.data_ptr(), flat.numel() * flat.element_size(), 0)
            self._pinned = int(rc) == 0
        if multi:
            dist.barrier(group=group)

    def close(self) -> None:
        if self._flat is None:
            return
        if self._pinned:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self._flat.data_ptr())
            self._pinned = False
        if self._multi:
            dist.barrier(group=self.group)
        self.tensor = self._flat = None
        if self.rank == 0 and os.path.exists(self.path):
            os.remove(self.path)

    def __del__(self):
        try:
            if self._flat is not None and self._pinned:
                torch.cuda.cudart().cudaHostUnregister(self._flat.data_ptr())
        except Exception:
            pass
