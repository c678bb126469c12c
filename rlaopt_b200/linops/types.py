"""Type alias + runtime check for "a linear operator or a tensor" (``rlaopt/linops/types.py:22-38``)."""
from __future__ import annotations

from typing import Any, Union

import torch

from .base import _BaseLinOp

__all__ = ["LinOpType", "_is_linop_or_torch_tensor"]

LinOpType = Union[
    "LinOp",
    "TwoSidedLinOp",
    "SymmetricLinOp",
    "DistributedLinOp",
    "DistributedTwoSidedLinOp",
    "DistributedSymmetricLinOp",
]


def _is_linop_or_torch_tensor(param: Any, param_name: str) -> None:
    if not isinstance(param, (_BaseLinOp, torch.Tensor)):
        raise TypeError(
            f"{param_name} is of type {type(param).__name__}, "
            "but expected type LinOpType or torch.Tensor"
        )
