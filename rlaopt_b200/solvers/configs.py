"""Import-path alias of ``rlaopt/solvers/configs.py`` (definitions live in ``_configs.py``)."""
from ._configs import (PCGConfig, SAPAccelConfig, SAPConfig, SolverConfig, _get_solver_name,  # noqa: F401
                       _is_solver_config)
