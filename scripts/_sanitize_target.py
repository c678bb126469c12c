"""Small end-to-end exercise of every kernel added in round 2 (compute-sanitizer target; sizes kept tiny)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import ops
from rlaopt_b200.kernels import KernelConfig, LaplaceLinOp, Matern52LinOp, RBFLinOp
from rlaopt_b200.linops import apply_fused
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n, m, d = 300, 700, 9
A1, A2 = (torch.randn(n, d, generator=g) + 1.0).to(dev), (torch.randn(m, d, generator=g) + 1.0).to(dev)
for k in (1, 3, 8, 16, 70, 130):
    V = torch.randn(m, k, generator=g).to(dev)
    for cls in (RBFLinOp, Matern52LinOp, LaplaceLinOp):
        op = cls(A1, A2, KernelConfig(lengthscale=2.0, const_scaling=1.5))
        Y = op @ V
        W = torch.randn(n, k, generator=g).to(dev)
        Z = op.T @ W
        blk = torch.tensor([0, -1, 5, 17, 299])
        R = op.row_oracle(blk) @ V
        if k <= 64:
            Yf, G, S = apply_fused(op, V, alpha=-1.0, addend=W, beta=0.5, rhs=W, gamma=1.0, gram_with=W[:, : min(k, 5)].contiguous(), want_sqnorm=True)
            Yg, _, _ = apply_fused(op.row_oracle(blk), V, addend=W, beta=0.1, addend_idx=blk, rhs=W, gamma=-1.0, rhs_idx=blk)
os.environ["RLAOPT_B200_TC_CG2"] = "1"
X = torch.randn(520, 24, generator=g).to(dev) / 5
V = torch.randn(520, 130, generator=g).to(dev)
Y = RBFLinOp(X, X, KernelConfig(lengthscale=1.0)) @ V
torch.cuda.synchronize()
print("sanitize target done", float(Y.abs().sum()))
