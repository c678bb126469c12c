"""Random draws for sketches, power iterations and block sampling.

The reference draws Gaussian test matrices directly on ``solver_config.device``
(``rlaopt/sketches/gauss.py:46``, ``ortho.py:50``,
``spectral_estimators/spectral_norm.py:16``), so a CPU run and a CUDA run of the
same seeded solve see different random numbers.  ``host_rng()`` (or
``RLAOPT_B200_HOST_RNG=1``) makes every draw come from the seeded **CPU**
generator and then moves it to the device: the shapes and the order of the draws
are the reference's, so a seeded run on the GPU consumes exactly the random stream
the reference consumes on the CPU (this is what the solver parity tests use).
"""
from __future__ import annotations

import contextlib
import os

import torch

__all__ = ["randn", "host_rng", "host_rng_enabled", "replicated_rng", "sync_from_rank0"]

_HOST_RNG = [os.environ.get("RLAOPT_B200_HOST_RNG", "0") not in ("", "0", "false", "False")]


def host_rng_enabled() -> bool:
    return _HOST_RNG[0]


@contextlib.contextmanager
def host_rng(enabled: bool = True):
    """Context manager: draw on the CPU generator, then move to the target device."""
    prev = _HOST_RNG[0]
    _HOST_RNG[0] = bool(enabled)
    try:
        yield
    finally:
        _HOST_RNG[0] = prev


_REPLICATED = [None]  # (process group,) while replicated_rng() is active


@contextlib.contextmanager
def replicated_rng(group=None):
    """SPMD solves (one process per GPU, replicated solver state): every random draw -- sketch matrices,
    power-iteration starts, SAP coordinate blocks -- is taken from rank 0 and broadcast, so all ranks advance
    identical iterates whatever their local seeds are."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("replicated_rng needs an initialised torch.distributed process group")
    prev = _REPLICATED[0]
    _REPLICATED[0] = (group,)
    try:
        yield
    finally:
        _REPLICATED[0] = prev


def sync_from_rank0(t: torch.Tensor, device: torch.device | None = None) -> torch.Tensor:
    """Broadcast ``t`` from rank 0 of the active ``replicated_rng`` group (identity otherwise).  Host tensors are
    staged through ``device`` when the backend cannot broadcast CPU memory (NCCL)."""
    if _REPLICATED[0] is None:
        return t
    import torch.distributed as dist

    (group,) = _REPLICATED[0]
    if dist.get_world_size(group) == 1:
        return t
    src = dist.get_global_rank(group, 0) if group is not None else 0
    if t.device.type == "cpu" and dist.get_backend(group) == "nccl":
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        staged = t.to(device)
        dist.broadcast(staged, src=src, group=group)
        return staged.cpu()
    t = t.contiguous()
    dist.broadcast(t, src=src, group=group)
    return t


def randn(*shape: int, dtype: torch.dtype | None = None, device: torch.device | str | None = None) -> torch.Tensor:
    """``torch.randn`` on ``device``; under ``host_rng`` the numbers come from the CPU stream."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if _HOST_RNG[0] and device.type != "cpu":
        out = torch.randn(*shape, dtype=dtype).to(device, non_blocking=True)
    else:
        out = torch.randn(*shape, dtype=dtype, device=device)
    return sync_from_rank0(out)
