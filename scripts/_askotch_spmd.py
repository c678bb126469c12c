"""torchrun worker: time ASkotch steps of the C4 shape (RBF, n = 10M, d = 16, blk = n/100, Nystrom rank 100) with the
SPMD kernel operator (row oracle column-sharded + all-reduce, block oracle row-sharded + all-gather)."""
import os, sys, time, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from rlaopt_b200.kernels import KernelConfig
from rlaopt_b200.kernels.sharded import replicate_from_host, sharded_kernel_linop
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.solvers import SAPAccelConfig, SAPConfig
from rlaopt_b200.solvers.sap import SAP
from rlaopt_b200.utils import replicated_rng
n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 16, 1
g = torch.Generator().manual_seed(0)
X = torch.randn(n, d, generator=g) / d**0.5
B = torch.randn(n, k, generator=g)
Xd = replicate_from_host(X.pin_memory(), dev)
A = sharded_kernel_linop(Xd, Xd, KernelConfig(lengthscale=1.0), "rbf", dev)
reg = 1e-2
system = LinSys(A, B.to(dev), reg=reg, A_row_oracle=A.row_oracle, A_blk_oracle=A.blk_oracle)
solver = SAP(system=system, W_init=torch.zeros(n, k, device=dev), precond_config=NystromConfig(rank=100, rho=reg), device=dev,
             blk_sz=n // 100, accel=True, accel_config=SAPAccelConfig(mu=reg, nu=100.0), power_iters=10)
solver.block_sampler = "device"
with replicated_rng():
    for _ in range(3): solver._step()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(10): solver._step()
    torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 10
    rel = system._compute_internal_metrics(solver.W)["rel_res"]
if rank == 0:
    print(f"SPMD ASkotch n={n} d={d} blk={n//100} rank=100 on {world} GPUs: {dt*1e3:.1f} ms per step; rel_res after 13 steps {float(rel.max()):.4f}", flush=True)
dist.destroy_process_group()
