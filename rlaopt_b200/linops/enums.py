"""Enums of the distribution protocol (interface of ``rlaopt/linops/enums.py:4-29``)."""
from enum import Enum, auto


class _Operation(Enum):
    MATVEC = auto()
    RMATVEC = auto()


class _DistributionMode(Enum):
    ROW = auto()  # shards own blocks of rows
    COLUMN = auto()  # shards own blocks of columns

    @classmethod
    def _from_str(cls, value, param_name):
        if isinstance(value, cls):
            return value
        lookup = {"row": cls.ROW, "column": cls.COLUMN}
        if isinstance(value, str) and value.lower() in lookup:
            return lookup[value.lower()]
        raise ValueError(
            f"Invalid value for {param_name}: {value}. "
            "Expected 'row', 'column', _DistributionMode.ROW, "
            "or _DistributionMode.COLUMN."
        )

    def flipped(self) -> "_DistributionMode":
        return _DistributionMode.COLUMN if self is _DistributionMode.ROW else _DistributionMode.ROW
