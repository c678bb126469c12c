"""A/B of the cta_group::2 pair kernels of the k > 64 family (run once per RLAOPT_B200_TC_CG2 setting)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig

dev = torch.device("cuda:0")
CASES = [("RBFLinOp", 94720, 2_000_000, 64, 1000), ("RBFLinOp", 37888, 1_000_000, 64, 128), ("RBFLinOp", 100000, 100000, 16, 100),
         ("Matern52LinOp", 37888, 1_000_000, 32, 200), ("RBFLinOp", 37888, 1_000_000, 128, 256)]
print("cg2:", os.environ.get("RLAOPT_B200_TC_CG2", "1"))
for name, n, m, d, k in CASES:
    g = torch.Generator(device=dev).manual_seed(0)
    A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
    A1 = A2[:n]
    V = torch.randn(m, k, generator=g, device=dev)
    op = getattr(K, name)(A1, A2, KernelConfig(lengthscale=1.0))
    Y = op @ V
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); Y = op @ V; b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[1]
    rows = torch.arange(0, n, max(n // 64, 1), device=dev)
    D2 = torch.cdist(A1[rows].double(), A2.double()).pow(2)
    if name == "RBFLinOp":
        Kr = torch.exp(-0.5 * D2)
    else:
        s5 = (5.0 * D2).sqrt(); Kr = (1 + s5 + s5 * s5 / 3) * torch.exp(-s5)
    ref = Kr @ V.double()
    err = float((Y[rows].double() - ref).norm() / ref.norm())
    print(f"{name:14s} n={n} m={m} d={d} k={k}: {t:8.2f} ms {n * m / t / 1e6:8.1f} Gentries/s  rel err {err:.2e}", flush=True)
