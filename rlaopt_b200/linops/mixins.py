"""Output scaling for linear operators (interface of ``rlaopt/linops/mixins.py:8-72``).

The kernel operators fold ``const_scaling`` into the CUDA epilogue instead of a
second elementwise pass; the mixin is kept for operators built from plain
callables and for attribute compatibility (``_scaling``).
"""
from __future__ import annotations

import functools

__all__ = ["ScaleMixin"]


class _ScaledFunction:
    """``scale * fn(...)`` as a picklable callable that remembers both parts."""

    def __init__(self, fn, scale: float):
        self.fn, self.scale = fn, scale
        try:
            functools.update_wrapper(self, fn)
        except (AttributeError, TypeError):
            pass
        name = getattr(fn, "__name__", None)
        if name:
            self.__name__ = f"scaled_{name}"

    def __call__(self, *args, **kwargs):
        return self.scale * self.fn(*args, **kwargs)


class ScaleMixin:
    """Adds a constant output scale to an operator."""

    def _initialize_scaling(self, scale) -> None:
        self._scaling = 1.0 if scale is None else float(scale)

    def _apply_scaling(self, obj):
        """Scale a result, or wrap a callable so that its results are scaled.

        A scale of exactly 1.0 returns ``obj`` untouched (``mixins.py:60-61``);
        wrapping an already scaled callable multiplies the scales (``:66-67``).
        """
        scale = getattr(self, "_scaling", 1.0)
        if scale == 1.0:
            return obj
        if isinstance(obj, _ScaledFunction):
            return _ScaledFunction(obj.fn, scale * obj.scale)
        if callable(obj):
            return _ScaledFunction(obj, scale)
        return scale * obj
