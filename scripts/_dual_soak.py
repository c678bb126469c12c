"""Soak test: the two-chunk kernels must reproduce the one-chunk kernel bit for bit (same arithmetic and summation order).
mode = "<RLAOPT_B200_TC_DUAL>[:<RLAOPT_B200_TC_DUAL_OVERLAP>]"."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig

dev = torch.device("cuda:0")
def run(op, V, mode):
    f = mode.split(":")
    os.environ["RLAOPT_B200_TC_DUAL"] = f[0]
    if len(f) > 1:
        os.environ["RLAOPT_B200_TC_DUAL_OVERLAP"] = f[1]
    else:
        os.environ.pop("RLAOPT_B200_TC_DUAL_OVERLAP", None)
    Y = op @ V
    torch.cuda.synchronize()
    return Y
CASES = [("RBFLinOp", 37888, 1_000_000, 128, 256, [("1", 6), ("2", 2)]),
         ("Matern32LinOp", 37888, 500_000, 100, 1000, [("1", 4)]),
         ("RBFLinOp", 94720, 2_000_000, 64, 1000, [("3", 8), ("2", 2)]),
         ("Matern52LinOp", 37888, 1_000_000, 32, 200, [("3", 6), ("2", 2)]),
         ("RBFLinOp", 50000, 300_000, 16, 512, [("3", 6)])]
for name, n, m, d, k, modes in CASES:
    g = torch.Generator(device=dev).manual_seed(1)
    A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
    V = torch.randn(m, k, generator=g, device=dev)
    op = getattr(K, name)(A2[:n], A2, KernelConfig(lengthscale=1.0))
    Y0 = run(op, V, "0")
    for mode, reps in modes:
        bad = [int((run(op, V, mode) != Y0).sum()) for _ in range(reps)]
        print(f"{name} n={n} m={m} d={d} k={k} DUAL={mode}: mismatching entries per launch {bad}", flush=True)
    del op, A2, V, Y0
    torch.cuda.empty_cache()
