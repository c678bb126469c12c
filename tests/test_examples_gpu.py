"""The example scripts (drop-in versions of the reference's experiment flow) run end to end on the visible GPUs."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_krr_pcg_example():
    import krr_pcg

    W, log = krr_pcg.main(n=6000, d=8, k=2, rank=100)
    assert bool((log[max(log)]["metrics"]["internal_metrics"]["rel_res"] <= 1e-4).all())


def test_distributed_askotch_example():
    """Reference flow: DistributedRBFLinOp over a *set* of devices + LinSys with its row / block oracles + accelerated
    SAP with a Nystrom block preconditioner (experiments/distributed_krr_linsys_askotch_solve_test.py)."""
    import distributed_krr_askotch as ex

    devices = [torch.device("cuda", i) for i in range(min(2, torch.cuda.device_count()))]
    W, log = ex.main(n=20000, d=10, k=3, max_iters=60, callback_freq=20, devices=devices)
    rel = [float(log[i]["metrics"]["internal_metrics"]["rel_res"].max()) for i in sorted(log)]
    # with the reference's parameters (mu = reg, nu = 100, 1 % blocks) ASkotch needs thousands of steps and its residual
    # is not monotone at the start (same iterates as the reference, test_solvers_gpu.py); here: the flow runs and stays bounded
    assert sorted(log) == [0, 20, 40, 60] and all(r == r and r < 10.0 for r in rel)
    assert W.shape == (20000, 3) and W.device == devices[0] and bool(torch.isfinite(W).all())
