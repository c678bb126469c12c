"""A/B of the two-chunk kernels (k > 128): RLAOPT_B200_TC_DUAL = 0 (one chunk), 2 (two chunks, drains behind the pointwise
stage), 3 (two chunks, pointwise stage sliced between the drains; d <= 64).  The knob is read per launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig

dev = torch.device("cuda:0")
CASES = [("RBFLinOp", 1000, 777, 64, 130), ("RBFLinOp", 4096, 8192 + 64, 20, 1000), ("Matern32LinOp", 3000, 64 * 5, 64, 257),
         ("RBFLinOp", 94720, 2_000_000, 64, 1000), ("Matern52LinOp", 37888, 1_000_000, 32, 200),
         ("RBFLinOp", 100000, 100000, 16, 300), ("RBFLinOp", 37888, 1_000_000, 128, 256)]
MODES = sys.argv[1:] or ["0", "2", "3n", "3"]
for name, n, m, d, k in CASES:
    g = torch.Generator(device=dev).manual_seed(0)
    A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
    A1 = A2[:n] if n <= m else torch.randn(n, d, generator=g, device=dev) / d**0.5
    V = torch.randn(m, k, generator=g, device=dev)
    op = getattr(K, name)(A1, A2, KernelConfig(lengthscale=1.0))
    rows = torch.arange(0, n, max(n // 64, 1), device=dev)
    D2 = torch.cdist(A1[rows].double(), A2.double()).pow(2)
    if name == "RBFLinOp":
        Kr = torch.exp(-0.5 * D2)
    elif name == "Matern32LinOp":
        s3 = (3.0 * D2).sqrt(); Kr = (1 + s3) * torch.exp(-s3)
    else:
        s5 = (5.0 * D2).sqrt(); Kr = (1 + s5 + s5 * s5 / 3) * torch.exp(-s5)
    ref = Kr @ V.double()
    for mode in MODES:
        os.environ["RLAOPT_B200_TC_DUAL"] = mode[0]
        os.environ["RLAOPT_B200_TC_DUAL_OVERLAP"] = "0" if mode.endswith("n") else "1"  # "3n": sliced, drains and quarters not overlapped
        Y = op @ V
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); Y = op @ V; b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts)[1]
        err = float((Y[rows].double() - ref).norm() / ref.norm())
        print(f"DUAL={mode} {name:14s} n={n} m={m} d={d} k={k}: {t:8.2f} ms {n * m / t / 1e6:8.1f} Gentries/s  rel err {err:.2e}", flush=True)
