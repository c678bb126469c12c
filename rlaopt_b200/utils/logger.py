"""Iteration logger used by ``Model._train`` (interface of ``rlaopt/utils/logger.py:11-51``).

``_compute_log(i, W)`` evaluates ``log_fn`` every ``log_freq`` iterations and
returns ``{"iter_time", "cum_time", "metrics"}`` (``None`` otherwise).  Times
exclude the metric evaluation itself, as in the reference; on CUDA the device is
synchronised before the clock is read so that asynchronous kernel launches are
charged to the iteration that issued them.  ``wandb`` is imported only when a run
is actually requested.
"""
from __future__ import annotations

import time
from typing import Callable

import torch

__all__ = ["Logger"]


def _now() -> float:
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        torch.cuda.synchronize()
    return time.time()


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
        self.start_time = _now()

    def _reset_timer(self):
        self.start_time = _now()

    def _update_cum_time(self):
        self.iter_time = _now() - self.start_time
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
