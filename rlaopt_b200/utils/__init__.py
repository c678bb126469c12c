"""Host-side helpers (validators, logger, reproducible random draws)."""
from .input_checkers import *  # noqa: F401,F403
from . import input_checkers as _ic
from .logger import Logger
from .rng import host_rng, host_rng_enabled, randn, replicated_rng, sync_from_rank0
from .shared_host import SharedPinnedTensor, shared_host_available

__all__ = list(_ic.__all__) + ["Logger", "host_rng", "host_rng_enabled", "randn", "replicated_rng", "sync_from_rank0",
                                 "SharedPinnedTensor", "shared_host_available"]
