"""Linear-operator interface of the kernel-matmat path (mirror of ``rlaopt.linops``)."""
from . import distributed, fused, mixins, simple, types
from .distributed import *  # noqa: F401,F403
from .fused import *  # noqa: F401,F403
from .mixins import *  # noqa: F401,F403
from .simple import *  # noqa: F401,F403
from .types import *  # noqa: F401,F403

__all__ = [name for mod in (distributed, fused, mixins, simple, types) for name in getattr(mod, "__all__", [])]
