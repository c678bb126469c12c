"""Throughput of the CUDA-core kernel family on one GPU (run once per library build, see RLAOPT_B200_LIB).

    RLAOPT_B200_LAYOUT=simt python scripts/_simt_ab.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RLAOPT_B200_LAYOUT", "simt")
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig

dev = torch.device("cuda:0")
CASES = [  # (class, dtype, n, m, d, k)
    ("LaplaceLinOp", torch.float32, 113664, 1 << 21, 32, 16),
    ("LaplaceLinOp", torch.float32, 65536, 1 << 20, 32, 16),
    ("LaplaceLinOp", torch.float32, 65536, 1 << 20, 32, 8),
    ("LaplaceLinOp", torch.float32, 65536, 1 << 20, 8, 10),
    ("LaplaceLinOp", torch.float32, 32768, 1 << 19, 128, 64),
    ("LaplaceLinOp", torch.float32, 65536, 1 << 20, 16, 1),
    ("RBFLinOp", torch.float32, 65536, 1 << 20, 32, 16),
    ("Matern52LinOp", torch.float32, 32768, 1 << 19, 128, 64),
    ("LaplaceLinOp", torch.float64, 16384, 1 << 18, 32, 16),
]
print("lib:", os.environ.get("RLAOPT_B200_LIB", "default"), "regp:", os.environ.get("RLAOPT_B200_SIMT_REGP", "1"))
for name, dt, n, m, d, k in CASES:
    g = torch.Generator().manual_seed(0)
    A1 = (torch.randn(n, d, generator=g, dtype=dt) / d**0.5).to(dev)
    A2 = (torch.randn(m, d, generator=g, dtype=dt) / d**0.5).to(dev)
    V = torch.randn(m, k, generator=g, dtype=dt).to(dev)
    op = getattr(K, name)(A1, A2, KernelConfig(lengthscale=1.0))
    Y = op @ V
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Y = op @ V
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[1]
    print(f"{name:14s} {str(dt)[6:]:8s} n={n} m={m} d={d} k={k}: {t:8.2f} ms  {n * m / t / 1e6:8.1f} Gentries/s  "
          f"checksum {Y.double().abs().sum().item():.10e}")
