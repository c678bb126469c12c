"""Iteration logger used by ``Model._train`` (interface of ``rlaopt/utils/logger.py:11-51``).

``_compute_log(i, W)`` evaluates ``log_fn`` every ``log_freq`` iterations and
returns ``{"iter_time", "cum_time", "metrics"}`` (``None`` otherwise).  Times
exclude the metric evaluation itself, as in the reference; on CUDA the interval is
measured with events on the stream (device time of the iterations' own work), the
reference reads ``time.time()`` without synchronising (``utils/logger.py:28-30``).  ``wandb`` is imported only when a run
is actually requested.
"""
from __future__ import annotations

import time
from typing import Callable

import torch

__all__ = ["Logger"]


class _Clock:
    """Elapsed time between two marks.  With a CUDA context the marks are events on the current stream
    (``torch.cuda.Event``): the interval is the device time of the work enqueued between them, and reading it waits
    for the closing event only -- not ``cuda.synchronize()`` + wall clock.  Without CUDA it is ``time.perf_counter``."""

    def __init__(self):
        self.cuda = torch.cuda.is_available() and torch.cuda.is_initialized()
        self.mark()

    def mark(self) -> None:
        if self.cuda:
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t0.record()
        else:
            self._t0 = time.perf_counter()

    def elapsed(self) -> float:
        """Seconds since the last mark."""
        if not self.cuda:
            return time.perf_counter() - self._t0
        t1 = torch.cuda.Event(enable_timing=True)
        t1.record()
        t1.synchronize()
        return self._t0.elapsed_time(t1) * 1e-3


class Logger:
    def __init__(self, log_freq: int, log_fn: Callable, wandb_kwargs: dict | None):
        self.log_freq = log_freq
        self.log_fn = log_fn
        self.log_in_wandb = wandb_kwargs is not None
        self._wandb = None
        if self.log_in_wandb:
            import wandb  # deferred: optional dependency

            self._wandb = wandb
            wandb.init(**wandb_kwargs)
        self.iter_time = 0
        self.cum_time = 0
        self._clock = _Clock()

    def _reset_timer(self):
        self._clock.mark()

    def _update_cum_time(self):
        self.iter_time = self._clock.elapsed()
        self.cum_time += self.iter_time

    def _compute_log(self, i: int, *args, **kwargs):
        if i % self.log_freq:
            return None
        self._update_cum_time()
        log_dict = {"iter_time": self.iter_time, "cum_time": self.cum_time, "metrics": self.log_fn(*args, **kwargs)}
        if self.log_in_wandb:
            self._wandb.log(log_dict, step=i)
        self._reset_timer()
        return log_dict

    def _terminate(self):
        if self.log_in_wandb:
            self._wandb.finish()
