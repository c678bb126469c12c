"""Kernel ridge regression with distributed ASkotch -- the flow of the reference's
``experiments/distributed_krr_linsys_askotch_solve_test.py`` with ``rlaopt`` replaced by ``rlaopt_b200``.

    python examples/distributed_krr_askotch.py [n] [d] [k] [max_iters]

Single process; the kernel operator spreads its row blocks over every visible GPU (``devices=set(...)``), the row
oracle runs column-distributed and the block oracle row-distributed, exactly as in the reference.  For the
one-process-per-GPU form see ``scripts/_askotch_spmd.py`` (``torchrun``, ``sharded_kernel_linop``).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlaopt_b200.kernels import DistributedRBFLinOp, KernelConfig  # noqa: E402
from rlaopt_b200.models import LinSys  # noqa: E402
from rlaopt_b200.preconditioners import NystromConfig  # noqa: E402
from rlaopt_b200.solvers import SAPAccelConfig, SAPConfig  # noqa: E402


def main(n=1_000_000, d=50, k=10, max_iters=300, callback_freq=100, devices=None):
    dtype = torch.float32
    torch.manual_seed(0)
    sigma, reg = 1.0, 1e-2
    if devices is None:
        devices = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]

    # synthetic data
    A = torch.randn(n, d, device=devices[0], dtype=dtype) / d**0.5
    b = torch.randn(n, k, device=devices[0], dtype=dtype)

    # linear operator for the kernel matrix
    lin_op = DistributedRBFLinOp(A1=A, A2=A, kernel_config=KernelConfig(lengthscale=sigma), devices=set(devices))
    try:
        system = LinSys(A=lin_op, B=b, reg=reg, A_row_oracle=lin_op.row_oracle, A_blk_oracle=lin_op.blk_oracle)
        solver_config = SAPConfig(
            precond_config=NystromConfig(rank=100, rho=reg),
            max_iters=max_iters,
            atol=1e-6,
            rtol=1e-6,
            blk_sz=max(n // 100, 1),
            accel_config=SAPAccelConfig(mu=reg, nu=100.0),
            device=devices[0],
        )
        W, log = system.solve(solver_config=solver_config, W_init=torch.zeros(n, k, device=devices[0], dtype=dtype),
                              callback_freq=callback_freq)
    finally:
        lin_op.shutdown()
    for it in sorted(log):
        rel = log[it]["metrics"]["internal_metrics"]["rel_res"]
        print(f"iter {it:5d}  cum_time {log[it]['cum_time']:8.3f} s  max rel_res {float(rel.max()):.4e}")
    return W, log


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:5]]
    main(*args)
