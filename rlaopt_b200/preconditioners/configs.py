"""Import-path alias of ``rlaopt/preconditioners/configs.py`` (definitions live in ``_configs.py``)."""
from ._configs import *  # noqa: F401,F403
from ._configs import (IdentityConfig, NewtonConfig, NystromConfig, PreconditionerConfig, SkPreConfig,  # noqa: F401
                       _is_precond_config)
