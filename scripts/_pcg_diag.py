"""Diagnose the C5-shaped solve (Nystrom rank r + block PCG with 16 right-hand sides) on one GPU:
matmat accuracy at k = r and k = 16 against fp64 sampled rows, Nystrom factor sanity, residual per iteration."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, RBFLinOp
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.preconditioners._precond import Nystrom
from rlaopt_b200.solvers import PCGConfig

dev = torch.device("cuda:0")
n, rank, reg = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
d, k = 64, 16
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(n, d, generator=g, device=dev) / d**0.5
B = torch.randn(n, k, generator=g, device=dev)
A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
rows = torch.randperm(n, device=dev)[:256]
Xd = X.double()
Kr = torch.exp(-0.5 * torch.cdist(Xd[rows], Xd).pow(2))
for kk in (rank, k, 1):
    V = torch.randn(n, kk, generator=g, device=dev)
    Y = A @ V
    ref = Kr @ V.double()
    print(f"matmat k={kk}: rel err {float((Y[rows].double() - ref).norm() / ref.norm()):.3e}  finite {bool(torch.isfinite(Y).all())}")
torch.manual_seed(0)
P = Nystrom(NystromConfig(rank=rank, rho=reg, sketch="gauss"))
P._update(A, dev)
print("S top", P.S[:4].tolist(), "S tail", P.S[-4:].tolist(), "finite U", bool(torch.isfinite(P.U).all()))
G = P.U.T @ P.U
print("||U^T U - I||_max", float((G - torch.eye(rank, device=dev)).abs().max()))
P._update_damping(baseline_rho=reg)
print("rho", float(P.config.rho))
x = torch.randn(n, k, generator=g, device=dev)
y = P._inv @ (P @ x)
print("||P^-1 P x - x|| / ||x||", float((y - x).norm() / x.norm()))
system = LinSys(A, B, reg=reg)
cfg = PCGConfig(device=dev, max_iters=40, rtol=1e-4, atol=1e-30, precond_config=NystromConfig(rank=rank, rho=reg, sketch="gauss"))
torch.manual_seed(0)
try:
    W, log = system.solve(cfg, torch.zeros(n, k, device=dev), callback_freq=1)
    for it in sorted(log):
        r = log[it]["metrics"]["internal_metrics"]["rel_res"]
        print(f"  iter {it:3d} rel_res max {float(r.max()):.3e} min {float(r.min()):.3e}")
except Exception as e:
    print("ERR", type(e).__name__, str(e)[:200])
