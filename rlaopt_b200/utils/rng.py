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

__all__ = ["randn", "host_rng", "host_rng_enabled"]

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


def randn(*shape: int, dtype: torch.dtype | None = None, device: torch.device | str | None = None) -> torch.Tensor:
    """``torch.randn`` on ``device``; under ``host_rng`` the numbers come from the CPU stream."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if _HOST_RNG[0] and device.type != "cpu":
        return torch.randn(*shape, dtype=dtype).to(device, non_blocking=True)
    return torch.randn(*shape, dtype=dtype, device=device)
