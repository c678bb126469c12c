import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for name, n, d, k in (("RBFLinOp", 32768, 32, 16), ("LaplaceLinOp", 32768, 32, 16), ("Matern52LinOp", 32768, 128, 64), ("RBFLinOp", 32768, 8, 1)):
    X = (torch.randn(n, d, generator=g, dtype=torch.float64) / d**0.5).to(dev)
    V = torch.randn(n, k, generator=g, dtype=torch.float64).to(dev)
    op = getattr(K, name)(X, X, KernelConfig(lengthscale=1.0))
    for _ in range(2): Y = op @ V
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): Y = op @ V
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"fp64 {name} n={n} d={d} k={k}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s  ({n*n*(2*d+2*k)/ms/1e9:.1f} TFLOP/s algorithmic)", flush=True)
