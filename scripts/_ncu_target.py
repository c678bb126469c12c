"""ncu target: a few products of one kernel family on one GPU.
    python scripts/_ncu_target.py <LinOpClass> <n> <m> <d> <k>"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig
dev = torch.device("cuda:0")
name, n, m, d, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
g = torch.Generator(device=dev).manual_seed(0)
A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
V = torch.randn(m, k, generator=g, device=dev)
if k == 1:
    V = V[:, 0].contiguous()
op = getattr(K, name)(A2[:n], A2, KernelConfig(lengthscale=1.0))
for _ in range(2):
    Y = op @ V
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); Y = op @ V; b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b)
print(f"{name} n={n} m={m} d={d} k={k}: {t:.2f} ms {n * m / t / 1e6:.1f} Gentries/s checksum {float(Y.double().abs().sum()):.6e}")
