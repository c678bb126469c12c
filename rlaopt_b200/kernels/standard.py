"""The ten public kernel operators (``rlaopt/kernels/standard.py:7-18,88-111``).

Formulas (``standard.py:31-85``), with ``u = (x - y) / lengthscale`` and ``r = |u|_2``:

    RBF       exp(-|u|_2^2 / 2)
    Laplace   exp(-|u|_1)
    Matern12  exp(-r)
    Matern32  (1 + sqrt(3) r) exp(-sqrt(3) r)
    Matern52  (1 + sqrt(5) r + 5/3 r^2) exp(-sqrt(5) r)

They are evaluated inside the fused CUDA kernels (``rlaopt_b200/csrc``); this
module only binds names to kernel ids.
"""
from .factory import _create_kernel_classes

__all__ = [
    "RBFLinOp",
    "DistributedRBFLinOp",
    "LaplaceLinOp",
    "DistributedLaplaceLinOp",
    "Matern12LinOp",
    "DistributedMatern12LinOp",
    "Matern32LinOp",
    "DistributedMatern32LinOp",
    "Matern52LinOp",
    "DistributedMatern52LinOp",
]

RBFLinOp, DistributedRBFLinOp = _create_kernel_classes("RBF", "rbf")
LaplaceLinOp, DistributedLaplaceLinOp = _create_kernel_classes("Laplace", "laplace")
Matern12LinOp, DistributedMatern12LinOp = _create_kernel_classes("Matern12", "matern12")
Matern32LinOp, DistributedMatern32LinOp = _create_kernel_classes("Matern32", "matern32")
Matern52LinOp, DistributedMatern52LinOp = _create_kernel_classes("Matern52", "matern52")
