"""Kernel hyper-parameters (interface of ``rlaopt/kernels/configs.py:11-68``)."""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Union

import torch

from rlaopt_b200.utils import _is_float

__all__ = ["KernelConfig"]


@dataclass(kw_only=True)
class KernelConfig:
    """``const_scaling * k(x, y; lengthscale)``.

    ``lengthscale`` is a python float, or a 1-D tensor with one entry per feature
    (ARD).  ``const_scaling`` must be a python float (ints are rejected, as in the
    reference's ``_is_float`` check, ``configs.py:49``).
    """

    const_scaling: float = 1.0
    lengthscale: Union[float, torch.Tensor]

    def __post_init__(self) -> None:
        _is_float(self.const_scaling, "const_scaling")
        ls = self.lengthscale
        if isinstance(ls, torch.Tensor):
            if ls.ndim != 1:
                raise ValueError(f"lengthscale has {ls.ndim} dimensions, but expected 1 dimension")
        elif not isinstance(ls, float):
            raise TypeError(
                f"lengthscale is of type {type(ls).__name__}, but expected type float or torch.Tensor"
            )

    def to_dict(self) -> dict:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    def to(self, device: torch.device) -> "KernelConfig":
        """Config with a tensor lengthscale moved to ``device`` (scalar configs return ``self``)."""
        if not isinstance(self.lengthscale, torch.Tensor):
            return self
        return KernelConfig(const_scaling=self.const_scaling, lengthscale=self.lengthscale.to(device))


def _is_kernel_config(param: Any, param_name: str) -> None:
    if not isinstance(param, KernelConfig):
        raise TypeError(
            f"{param_name} is of type {type(param).__name__}, but expected type KernelConfig"
        )
