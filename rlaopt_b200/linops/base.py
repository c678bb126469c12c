"""Linear-operator base classes.

Mirrors the interface of ``rlaopt/linops/base.py``: ``_BaseLinOp`` (``:11-111``)
with ``@`` / right-``@`` dispatch on the rank of the argument, and
``_BaseDistributedLinOp`` (``:114-291``).  The distributed base keeps the
reference's constructor signature and ROW / COLUMN protocol, but instead of one
``torch.multiprocessing`` worker process per device with CPU-staged pickled
queues it drives every device from the calling process: the shard operators only
enqueue CUDA work, so one host thread keeps all GPUs busy, operands move
device-to-device (NVLink P2P), and nothing round-trips through host memory.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Sequence

import torch

from .enums import _DistributionMode, _Operation
from rlaopt_b200.utils import _is_list, _is_torch_device, _is_torch_f32_f64, _is_torch_size

__all__: list[str] = []


class _BaseLinOp(ABC):
    """Shape / dtype / device bookkeeping plus the matmul protocol."""

    def __init__(self, device: torch.device, shape: torch.Size, dtype: torch.dtype):
        self._check_inputs_base(device, shape, dtype)
        self._device, self._shape, self._dtype = device, shape, dtype

    def _check_inputs_base(self, device: Any, shape: Any, dtype: Any) -> None:
        _is_torch_device(device, "device")
        _is_torch_size(shape, "shape")
        if len(shape) != 2:
            raise ValueError(f"shape must have two elements. Received {len(shape)}")
        if any((not isinstance(s, int)) or s <= 0 for s in shape):
            raise ValueError(f"shape must contain positive integers. Received {shape}")
        _is_torch_f32_f64(dtype, "dtype")

    # -- metadata ---------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def devices(self) -> list[torch.device]:
        """All devices the operator computes on (one entry unless distributed)."""
        return [self._device]

    @property
    def shape(self) -> torch.Size:
        return self._shape

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @property
    def T(self) -> "_BaseLinOp":
        raise NotImplementedError("This linear operator doesn't support transposition")

    # -- products ---------------------------------------------------------
    @abstractmethod
    def _matvec(self, x: torch.Tensor) -> torch.Tensor:
        ...

    @abstractmethod
    def _matmat(self, x: torch.Tensor) -> torch.Tensor:
        ...

    def _rmatvec(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("This linear operator doesn't support right matvec")

    def _rmatmat(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("This linear operator doesn't support right matmat")

    @staticmethod
    def _rank_of(x: torch.Tensor) -> int:
        if x.ndim not in (1, 2):
            raise ValueError(f"x must be a 1D or 2D tensor. Received {x.ndim}D tensor.")
        return x.ndim

    def __matmul__(self, x: torch.Tensor) -> torch.Tensor:
        return self._matvec(x) if self._rank_of(x) == 1 else self._matmat(x)

    def __rmatmul__(self, x: torch.Tensor) -> torch.Tensor:
        # x @ A  ==  (A^T x^T)^T
        return self._rmatvec(x) if self._rank_of(x) == 1 else self._rmatmat(x.T).T


class _BaseDistributedLinOp(_BaseLinOp):
    """A linear operator split into per-device shards (rows or columns).

    ``A`` holds one shard operator per device.  ROW mode: shard ``i`` owns a block
    of output rows (apply = broadcast ``x`` + concatenate); COLUMN mode: shard ``i``
    owns a block of input columns (apply = scatter ``x`` + sum).  This is the
    protocol of ``rlaopt/linops/base.py:231-276`` / ``linops/distributed.py:40-50``.

    The ``is_new`` / ``manager`` / ``result_queue`` / ``task_queues`` / ``workers``
    parameters of the reference constructor are accepted for source compatibility
    and ignored: there are no worker processes to share.
    """

    def __init__(
        self,
        shape: torch.Size,
        A: list[_BaseLinOp],
        distribution_mode: Any,
        is_new: bool = True,
        manager=None,
        result_queue=None,
        task_queues=None,
        workers=None,
    ):
        self._is_new = is_new
        if is_new:
            _is_list(A, "A")
            if not all(isinstance(op, _BaseLinOp) for op in A):
                raise ValueError("All elements in A must be instances of _BaseLinOp")
            if len(A) == 0:
                raise ValueError("A must contain at least one linear operator")
            if any(op.dtype != A[0].dtype for op in A):
                raise ValueError(
                    "All linear operators must have the same dtype. "
                    f"Received {', '.join(str(op.dtype) for op in A)}."
                )
        super().__init__(device=torch.device("cpu"), shape=shape, dtype=A[0].dtype)  # device is a placeholder
        self._A = A
        self._devices = [op.device for op in A]
        self._distribution_mode = _DistributionMode._from_str(distribution_mode, "distribution_mode")
        self._closed = False

    @property
    def device(self):
        raise AttributeError(
            "Distributed linear operators operate across multiple devices "
            "and don't have a single 'device'. "
            "Use the 'devices' property instead to get the list of all devices."
        )

    @property
    def devices(self) -> list[torch.device]:
        return self._devices

    # -- the distribution protocol -----------------------------------------
    def _split_sizes(self, by_dimension: int) -> list[int]:
        return [op.shape[by_dimension] for op in self._A]

    def _chunk_tensor(self, x: torch.Tensor, by_dimension: int) -> list[torch.Tensor]:
        """Slices of ``x`` matching each shard's extent along ``by_dimension`` of the shard shape."""
        pieces, start = [], 0
        for size in self._split_sizes(by_dimension):
            pieces.append(x[start : start + size])
            start += size
        return pieces

    def _run_shards(self, x: torch.Tensor, operation: _Operation, chunk: bool, by_dimension: int) -> list[torch.Tensor]:
        """Apply every shard to (its slice of) ``x`` on the shard's own device.

        All transfers and kernels are enqueued before any result is consumed, so the
        devices run concurrently.
        """
        if self._closed:
            raise RuntimeError("distributed linear operator has been shut down")
        inputs: Sequence[torch.Tensor] = self._chunk_tensor(x, by_dimension) if chunk else [x] * len(self._A)
        # all transfers first, then all products: a device-to-device copy is issued on the SOURCE device's stream, so
        # a copy enqueued after that device's product would wait for it and serialise the devices
        moved = [xi.to(dev, non_blocking=True) for dev, xi in zip(self._devices, inputs)]
        return [op @ xi if operation == _Operation.MATVEC else op.T @ xi for op, xi in zip(self._A, moved)]

    @staticmethod
    def _combine_results(results: list[torch.Tensor], concatenate: bool, device: torch.device) -> torch.Tensor:
        # A non-blocking copy to the host returns before the data has landed (torch stages it through pinned
        # memory); only device-to-device moves may be asynchronous -- they are ordered on the target's stream.
        async_ok = torch.device(device).type == "cuda"
        moved = [r.to(device, non_blocking=async_ok) for r in results]
        if concatenate:
            return torch.cat(moved, dim=0)
        out = moved[0].clone() if len(moved) > 1 else moved[0]
        for r in moved[1:]:
            out += r
        return out

    # -- lifecycle -----------------------------------------------------------
    def shutdown(self) -> None:
        """Kept for API compatibility (the reference stops its worker processes here,
        ``rlaopt/linops/base.py:278-288``).  Idempotent; only owners close."""
        if self._is_new:
            self._closed = True
