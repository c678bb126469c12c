"""Import-path alias of ``rlaopt/solvers/sap.py``."""
from ._sap import SAP, VALID_PRECONDS  # noqa: F401
