"""rlaopt_b200 — B200-native implicit kernel-matrix matmat behind the rlaopt API.

``rlaopt_b200.kernels`` and ``rlaopt_b200.linops`` mirror ``rlaopt.kernels`` /
``rlaopt.linops`` for the one hot path this package rebuilds:
``Y = c * K(A1_rows, A2) @ V`` for RBF / Laplace / Matern kernels, evaluated by
hand-written sm_100a CUDA kernels behind the C ABI of ``include/rlaopt_b200.h``.
"""
import torch  # noqa: F401  (device memory, streams, torch.distributed plumbing)

from . import ops  # registers torch.ops.rlaopt_b200.kernel_matmat  # noqa: F401

ops.load_torch_op()  # torch.ops.rlaopt.kernel_matmat (C++ TORCH_LIBRARY_FRAGMENT in csrc/torch_op.cpp), when built
from . import linops, kernels  # noqa: F401
from . import sketches, preconditioners, spectral_estimators, solvers, models  # noqa: F401  consumers of the path

__version__ = "0.1.0"
