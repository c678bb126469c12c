"""Argument validators shared by the kernel / linop classes.

Same names, messages and exception types as the reference's
``rlaopt/utils/input_checkers.py:27-147`` (callers and tests match on
``TypeError`` / ``ValueError``), implemented through one table-driven helper.
"""
from __future__ import annotations

from typing import Any

import torch

__all__ = [
    "_is_bool",
    "_is_callable",
    "_is_dict",
    "_is_float",
    "_is_int",
    "_is_list",
    "_is_set",
    "_is_str",
    "_is_torch_device",
    "_is_torch_dtype",
    "_is_torch_f32_f64",
    "_is_torch_size",
    "_is_torch_tensor",
    "_is_torch_tensor_1d_2d",
    "_is_nonneg_float",
    "_is_pos_float",
    "_is_pos_int",
]


def _expect(param: Any, param_name: str, types, label: str) -> None:
    if not isinstance(param, types):
        raise TypeError(
            f"{param_name} is of type {type(param).__name__}, but expected type {label}"
        )


def _make(types, label):
    def check(param: Any, param_name: str) -> None:
        _expect(param, param_name, types, label)

    check.__name__ = f"_is_{label.replace('.', '_')}"
    return check


_is_bool = _make(bool, "bool")
_is_dict = _make(dict, "dict")
_is_float = _make(float, "float")
_is_int = _make(int, "int")
_is_list = _make(list, "list")
_is_set = _make(set, "set")
_is_str = _make(str, "str")
_is_torch_device = _make(torch.device, "torch.device")
_is_torch_dtype = _make(torch.dtype, "torch.dtype")
_is_torch_size = _make(torch.Size, "torch.Size")
_is_torch_tensor = _make(torch.Tensor, "torch.Tensor")


def _is_callable(param: Any, param_name: str) -> None:
    if not callable(param):
        raise TypeError(
            f"{param_name} is of type {type(param).__name__}, but expected type callable"
        )


def _is_torch_f32_f64(param: Any, param_name: str) -> None:
    _is_torch_dtype(param, param_name)
    if param not in (torch.float32, torch.float64):
        raise ValueError(
            f"{param_name} is {param}, but expected torch.float32 or torch.float64"
        )


def _is_torch_tensor_1d_2d(param: Any, param_name: str) -> None:
    _is_torch_tensor(param, param_name)
    if param.ndim not in (1, 2):
        raise ValueError(
            f"{param_name} must be a 1D or 2D tensor. Received {param.ndim}D tensor."
        )


def _is_nonneg_float(param: Any, param_name: str) -> None:
    _is_float(param, param_name)
    if param < 0:
        raise ValueError(f"{param_name} must be non-negative, but received {param}")


def _is_pos_float(param: Any, param_name: str) -> None:
    _is_float(param, param_name)
    if param <= 0:
        raise ValueError(f"{param_name} must be positive, but received {param}")


def _is_pos_int(param: Any, param_name: str) -> None:
    _is_int(param, param_name)
    if param <= 0:
        raise ValueError(f"{param_name} must be positive, but received {param}")
