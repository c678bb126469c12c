import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, LaplaceLinOp
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n, d, k = 65536, 32, 16
X = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
V = torch.randn(n, k, generator=g).to(dev)
op = LaplaceLinOp(X, X, KernelConfig(lengthscale=1.0))
for _ in range(3):
    Y = op @ V
torch.cuda.synchronize()
print(float(Y.abs().sum()))
