import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, RBFLinOp
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.preconditioners.nystrom import Nystrom
from rlaopt_b200.solvers import PCGConfig
from rlaopt_b200.sketches import get_sketch
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n, d, k = 20000, 8, 1
X = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
B = torch.randn(n, k, generator=g).to(dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out
ms, A = t(lambda: RBFLinOp(X, X, KernelConfig(lengthscale=1.0))); print(f"op ctor {ms:.2f} ms")
Om = get_sketch("gauss", "right", 200, n, torch.float32, dev)
ms, Y = t(lambda: A @ Om.Omega_mat); print(f"sketch matmat k=200 {ms:.2f} ms")
ms, _ = t(lambda: A @ B); print(f"matmat k=1 {ms:.2f} ms")
core = Om.Omega_mat.T @ Y; core.diagonal().add_(1e-7 * torch.trace(core))
ms, C = t(lambda: torch.linalg.cholesky(core)); print(f"cholesky 200 {ms:.2f} ms")
ms, F = t(lambda: torch.linalg.solve_triangular(C.T, Y, upper=True, left=False)); print(f"trsm {ms:.2f} ms")
ms, QR = t(lambda: torch.linalg.qr(F, mode='reduced')); print(f"qr 20000x200 {ms:.2f} ms")
ms, _ = t(lambda: torch.linalg.svd(QR[1], full_matrices=False)); print(f"svd 200x200 {ms:.2f} ms")
ms, _ = t(lambda: torch.linalg.eigh(QR[1] @ QR[1].T)); print(f"eigh 200x200 {ms:.2f} ms")
def build():
    P = Nystrom(NystromConfig(rank=200, rho=1.0, sketch="gauss")); P._update(A, dev); P._update_damping(1.0); return P
ms, P = t(build); print(f"nystrom build total {ms:.2f} ms")
R = B.clone()
ms, _ = t(lambda: P._inv @ R); print(f"P^-1 apply {ms:.2f} ms")
def solve():
    torch.manual_seed(1)
    s = LinSys(A, B, reg=1.0)
    return s.solve(PCGConfig(device=dev, max_iters=100, rtol=1e-4, precond_config=NystromConfig(rank=200, rho=1.0, sketch="gauss")), torch.zeros(n, k, device=dev), callback_freq=1)
ms, out = t(solve, reps=3); print(f"full solve {ms:.2f} ms, iters {max(out[1])}")
