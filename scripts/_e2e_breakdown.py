"""Where the end-to-end step of bench.py spends its time under torchrun (N >= 2): per-phase wall times with a
device synchronize after every phase.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/_e2e_breakdown.py
"""
import os, sys, time, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from rlaopt_b200.kernels import KernelConfig
from rlaopt_b200.kernels.sharded import replicate_from_host, sharded_kernel_linop

n, d, k = 1_000_000, 128, 64
g = torch.Generator().manual_seed(0)
Xp = (torch.randn(n, d, generator=g) / d**0.5).pin_memory()
Vp = torch.randn(n, k, generator=g).pin_memory()
from rlaopt_b200 import ops
from rlaopt_b200.utils import SharedPinnedTensor

shared = SharedPinnedTensor("e2e_breakdown_Y", (n, k), torch.float32)
Yh = shared.tensor
cfg = KernelConfig(lengthscale=1.0)


def tick(label, t0, acc):
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    acc.setdefault(label, []).append((t1 - t0) * 1e3)
    return t1


acc = {}
for step in range(5):
    dist.barrier(); torch.cuda.synchronize(dev)
    t = time.perf_counter(); t_start = t
    Xg = replicate_from_host(Xp, dev); t = tick("replicate X (H2D 1/N + all-gather)", t, acc)
    op = sharded_kernel_linop(Xg, Xg, cfg, "rbf", dev); t = tick("operator construction", t, acc)
    Vg = replicate_from_host(Vp, dev); t = tick("replicate V (H2D 1/N + all-gather)", t, acc)
    c = op.local_op._cache.center(); t = tick("column means of X", t, acc)
    P1, P2 = op.local_op._cache.get(ops.LAYOUT_TC); t = tick("pack row block + pack X", t, acc)
    ok = op.local_op._cache.tc_ok(0); t = tick("norm statistics read-back", t, acc)
    Yl = op.local_op @ Vg; t = tick("local block product (V pack + kernel)", t, acc)
    Yh[op.lo:op.hi].copy_(Yl, non_blocking=True); t = tick("D2H of the row block (every rank, own link)", t, acc)
    dist.barrier(); t = tick("barrier", t, acc)
    del op, Xg, Vg, Yl, P1, P2
    t = tick("free", t, acc)
    acc.setdefault("whole step", []).append((t - t_start) * 1e3)
if rank == 0:
    for key, v in acc.items():
        print(f"{key:55s} " + "  ".join(f"{x:8.1f}" for x in v) + "  ms")
shared.close()
dist.destroy_process_group()
