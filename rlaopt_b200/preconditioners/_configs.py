"""Preconditioner hyper-parameter records.

Field names, defaults and validation follow ``rlaopt/preconditioners/configs.py:28-132`` and
``enums.py:4-31``: keyword-only mutable dataclasses (``NystromConfig.rho`` is *rewritten* by
adaptive damping, ``nystrom.py:150-152``), ``damping_mode`` given as ``"adaptive"`` /
``"non_adaptive"`` and stored as a ``_DampingMode`` member.
"""
from __future__ import annotations

import enum
from dataclasses import asdict, dataclass
from typing import Any

from rlaopt_b200.utils import _is_nonneg_float, _is_pos_int


class _DampingMode(enum.Enum):
    ADAPTIVE = "adaptive"
    NON_ADAPTIVE = "non_adaptive"

    @classmethod
    def _from_str(cls, value, param_name):
        if isinstance(value, cls):
            return value
        if isinstance(value, str):
            for member in cls:
                if member.value == value.lower():
                    return member
        raise ValueError(
            f"Invalid value for {param_name}: {value}. Expected 'adaptive', 'non_adaptive', "
            "_DampingMode.ADAPTIVE, or _DampingMode.NON_ADAPTIVE.")


@dataclass(kw_only=True)
class PreconditionerConfig:
    def to_dict(self) -> dict:
        return asdict(self)


@dataclass(kw_only=True)
class IdentityConfig(PreconditionerConfig):
    pass


@dataclass(kw_only=True)
class NewtonConfig(PreconditionerConfig):
    rho: float

    def __post_init__(self):
        _is_nonneg_float(self.rho, "rho")


@dataclass(kw_only=True)
class NystromConfig(PreconditionerConfig):
    rank: int
    rho: float
    sketch: str = "ortho"
    damping_mode: str = "adaptive"

    def __post_init__(self):
        _is_pos_int(self.rank, "rank")
        _is_nonneg_float(self.rho, "rho")
        self.damping_mode = _DampingMode._from_str(self.damping_mode, "damping_mode")


@dataclass(kw_only=True)
class SkPreConfig(PreconditionerConfig):
    sketch_size: int
    rho: float
    sketch: str = "sparse"

    def __post_init__(self):
        _is_pos_int(self.sketch_size, "sketch_size")
        _is_nonneg_float(self.rho, "rho")


def _is_precond_config(param: Any, param_name: str):
    if not isinstance(param, PreconditionerConfig):
        raise TypeError(
            f"{param_name} is of type {type(param).__name__}, but expected type PreconditionerConfig")
