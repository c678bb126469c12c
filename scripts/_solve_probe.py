"""Convergence probes for the two solve-level BASELINE configs on ONE GPU at reduced n (how many ASkotch steps /
PCG iterations the bench legs must budget for):  python scripts/_solve_probe.py askotch n max_iters | pcg n reg"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, RBFLinOp
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.solvers import PCGConfig, SAPAccelConfig, SAPConfig

dev = torch.device("cuda:0")
mode = sys.argv[1]
if mode == "askotch":
    n, max_iters = int(sys.argv[2]), int(sys.argv[3])
    d, k, reg = 16, 1, float(sys.argv[4]) if len(sys.argv) > 4 else 1e-2
    os.environ["RLAOPT_B200_SAP_SAMPLER"] = "device"
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(n, d, generator=g, device=dev) / d**0.5
    B = torch.randn(n, k, generator=g, device=dev)
    A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
    system = LinSys(A, B, reg=reg, A_row_oracle=A.row_oracle, A_blk_oracle=A.blk_oracle)
    cfg = SAPConfig(precond_config=NystromConfig(rank=100, rho=reg), max_iters=max_iters, atol=1e-30, rtol=1e-4,
                    blk_sz=n // 100, accel_config=SAPAccelConfig(mu=reg, nu=float(sys.argv[5]) if len(sys.argv) > 5 else 100.0), device=dev)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    W, log = system.solve(cfg, torch.zeros(n, k, device=dev), callback_freq=100)
    torch.cuda.synchronize()
    print(f"askotch n={n}: {time.perf_counter() - t0:.1f} s wall")
    for it in sorted(log):
        print(f"  iter {it:5d} cum {log[it]['cum_time']:7.2f} s rel_res {float(log[it]['metrics']['internal_metrics']['rel_res'].max()):.3e}")
else:
    n, reg = int(sys.argv[2]), float(sys.argv[3])
    d, k, rank = 64, 16, 1000
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(n, d, generator=g, device=dev) / d**0.5
    B = torch.randn(n, k, generator=g, device=dev)
    A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
    system = LinSys(A, B, reg=reg)
    cfg = PCGConfig(device=dev, max_iters=int(sys.argv[4]) if len(sys.argv) > 4 else 60, rtol=float(sys.argv[5]) if len(sys.argv) > 5 else 1e-3,
                    atol=1e-30, precond_config=NystromConfig(rank=rank, rho=reg, sketch="gauss"))
    torch.manual_seed(0)
    t0 = time.perf_counter()
    try:
        W, log = system.solve(cfg, torch.zeros(n, k, device=dev), callback_freq=1, residual="recurrence")
    except Exception as e:
        print("ERR", type(e).__name__, str(e)[:200]); sys.exit(0)
    torch.cuda.synchronize()
    print(f"pcg n={n} reg={reg}: {time.perf_counter() - t0:.1f} s wall, {max(log)} iterations")
    for it in sorted(log):
        r = log[it]['metrics']['internal_metrics']['rel_res']
        print(f"  iter {it:4d} cum {log[it]['cum_time']:7.2f} s rel_res max {float(r.max()):.3e} min {float(r.min()):.3e}")
    true = system._true_sq_residual(W).sqrt() / system._rhs_norms()
    print(f"  true rel_res max {float(true.max()):.3e}")
