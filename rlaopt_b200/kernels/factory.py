"""Builds the public ``<Name>LinOp`` / ``Distributed<Name>LinOp`` classes.

Same public surface as ``rlaopt/kernels/factory.py:9-79``: constructor signatures
``(A1, A2, kernel_config)`` and ``(A1, A2, kernel_config, devices, use_full_kernel=True)``.
Where the reference binds a symbolic KeOps formula, these classes bind the id of
the fused CUDA kernel.
"""
from __future__ import annotations

import torch

from .base import _DistributedKernelLinOp, _KernelLinOp
from .configs import KernelConfig


def _create_kernel_classes(kernel_name: str, kernel_key: str):
    """Return ``(KernelLinOp, DistributedKernelLinOp)`` classes for one kernel."""

    class KernelLinOp(_KernelLinOp):
        def __init__(self, A1: torch.Tensor, A2: torch.Tensor, kernel_config: KernelConfig):
            super().__init__(A1=A1, A2=A2, kernel_config=kernel_config, _kernel_key=kernel_key)

    class DistributedKernelLinOp(_DistributedKernelLinOp):
        def __init__(
            self,
            A1: torch.Tensor,
            A2: torch.Tensor,
            kernel_config: KernelConfig,
            devices: set[torch.device],
            use_full_kernel: bool = True,
        ):
            super().__init__(
                A1=A1,
                A2=A2,
                kernel_config=kernel_config,
                devices=devices,
                use_full_kernel=use_full_kernel,
                _kernel_key=kernel_key,
            )

    KernelLinOp.__name__ = KernelLinOp.__qualname__ = f"{kernel_name}LinOp"
    DistributedKernelLinOp.__name__ = DistributedKernelLinOp.__qualname__ = f"Distributed{kernel_name}LinOp"
    KernelLinOp.__doc__ = f"{kernel_name} kernel linear operator (fused sm_100a CUDA matmat)."
    DistributedKernelLinOp.__doc__ = (
        f"{kernel_name} kernel linear operator row-partitioned over several GPUs, "
        "with row and block oracles."
    )
    return KernelLinOp, DistributedKernelLinOp
