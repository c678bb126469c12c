import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, RBFLinOp
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.solvers import SAPConfig, SAPAccelConfig
from rlaopt_b200.solvers.sap import SAP
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 16, 1
X = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
B = torch.randn(n, k, generator=g).to(dev)
A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
reg = 1e-2
system = LinSys(A, B, reg=reg, A_row_oracle=A.row_oracle, A_blk_oracle=A.blk_oracle)
cfg = SAPConfig(device=dev, max_iters=100, rtol=1e-4, blk_sz=n // 100, precond_config=NystromConfig(rank=100, rho=reg),
                accel_config=SAPAccelConfig(mu=reg, nu=100.0))
import os
solver = SAP(system=system, W_init=torch.zeros(n, k, device=dev), precond_config=cfg.precond_config, device=dev,
             blk_sz=cfg.blk_sz, accel=True, accel_config=cfg.accel_config, power_iters=10)
def sync(): torch.cuda.synchronize()
for _ in range(3): solver._step()
sync(); t0 = time.perf_counter()
for _ in range(10): solver._step()
sync(); print(f"n={n}: {(time.perf_counter()-t0)/10*1e3:.2f} ms per ASkotch step (blk {cfg.blk_sz}, rank 100)")
# breakdown
blk = solver._get_blk()
def t(fn, reps=5):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    sync(); return (time.perf_counter() - t0) / reps * 1e3, out
ms, Abb = t(lambda: system.A_blk_oracle(blk)); print(f"  blk_oracle ctor {ms:.2f} ms")
ms, P = t(lambda: solver._get_precond(blk, Abb)); print(f"  block Nystrom build {ms:.2f} ms")
ms, _ = t(lambda: solver._get_stepsize(blk, P, Abb)); print(f"  step size (<=10 power iterations) {ms:.2f} ms")
rows = blk.to(dev)
ms, Ar = t(lambda: system.A_row_oracle(rows)); print(f"  row_oracle ctor {ms:.2f} ms")
ms, _ = t(lambda: Ar @ solver.Y); print(f"  row oracle matvec ({len(blk)} x {n}) {ms:.2f} ms  -> {len(blk)*n/ms/1e6:.0f} Gentries/s")
ms, _ = t(lambda: Abb @ solver.Y[rows]); print(f"  block matvec {ms:.2f} ms")
