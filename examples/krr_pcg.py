"""Kernel ridge regression with Nystrom-preconditioned PCG on one GPU (BASELINE configs[0] shape by default).

    python examples/krr_pcg.py [n] [d] [k] [rank]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlaopt_b200.kernels import KernelConfig, RBFLinOp  # noqa: E402
from rlaopt_b200.models import LinSys  # noqa: E402
from rlaopt_b200.preconditioners import NystromConfig  # noqa: E402
from rlaopt_b200.solvers import PCGConfig  # noqa: E402


def main(n=20_000, d=8, k=1, rank=200, reg=1.0, rtol=1e-4):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    X = torch.randn(n, d, device=dev) / d**0.5
    B = torch.randn(n, k, device=dev)
    K = RBFLinOp(A1=X, A2=X, kernel_config=KernelConfig(lengthscale=1.0))
    system = LinSys(A=K, B=B, reg=reg)
    config = PCGConfig(device=dev, max_iters=200, rtol=rtol, precond_config=NystromConfig(rank=rank, rho=reg, sketch="gauss"))
    W, log = system.solve(solver_config=config, W_init=torch.zeros(n, k, device=dev), callback_freq=1)
    last = max(log)
    print(f"converged in {last} iterations, {log[last]['cum_time']:.3f} s, "
          f"max rel_res {float(log[last]['metrics']['internal_metrics']['rel_res'].max()):.3e}")
    return W, log


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:5]])
