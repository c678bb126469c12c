"""torchrun worker for test_spmd_gpu.py: row-sharded fused kernel operator over NCCL + replicated-state solvers."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import kernel_oracle as ko
    from rlaopt_b200.kernels import KernelConfig, Matern52LinOp, RBFLinOp
    from rlaopt_b200.kernels.sharded import replicate_from_host, sharded_kernel_linop
    from rlaopt_b200.models import LinSys
    from rlaopt_b200.preconditioners import NystromConfig
    from rlaopt_b200.solvers import PCGConfig, SAPAccelConfig, SAPConfig
    from rlaopt_b200.utils import replicated_rng

    g = torch.Generator().manual_seed(0)
    n, d, k = 3001, 12, 3  # ragged row blocks
    X = torch.randn(n, d, generator=g) / d**0.5
    V = torch.randn(n, k, generator=g)
    B = torch.randn(n, k, generator=g)
    cfg = KernelConfig(lengthscale=1.1, const_scaling=1.5)
    Xd = replicate_from_host(X.pin_memory(), dev)
    assert torch.equal(Xd.cpu(), X)
    ok = True
    for name, cls in (("rbf", RBFLinOp), ("matern52", Matern52LinOp), ("laplace", None)):
        op = sharded_kernel_linop(Xd, Xd, cfg, name, dev)
        Y = op @ V.to(dev)
        ref = ko.kernel_matmat(X, X, V, name, 1.1, 1.5, dtype=torch.float64)
        ok &= ko.rel_fro_error(Y, ref) <= 1e-5
        Yt = op.T @ V.to(dev)
        ok &= ko.rel_fro_error(Yt, ref) <= 1e-5  # symmetric kernel matrix
        if cls is not None:  # bitwise equal to the single-GPU operator on the same rows
            single = cls(Xd, Xd, cfg) @ V.to(dev)
            ok &= bool(torch.allclose(Y, single, rtol=1e-6, atol=1e-6))
    # SPMD oracles: row oracle = column-sharded partial products + all-reduce, block oracle = row-sharded + all-gather
    A_sh = sharded_kernel_linop(Xd, Xd, cfg, "rbf", dev)
    A_one = RBFLinOp(Xd, Xd, cfg)
    gb = torch.Generator().manual_seed(3)
    for b in (1, 7, 301, 1000):
        blk = torch.randperm(n, generator=gb)[:b]
        Vb = torch.randn(b, k, generator=gb).to(dev)
        ok &= bool(torch.allclose(A_sh.row_oracle(blk) @ V.to(dev), A_one.row_oracle(blk.to(dev)) @ V.to(dev), rtol=2e-5, atol=2e-5))
        ok &= bool(torch.allclose(A_sh.blk_oracle(blk) @ Vb, A_one.blk_oracle(blk.to(dev)) @ Vb, rtol=2e-5, atol=2e-5))
        ok &= bool(torch.allclose(A_sh.blk_oracle(blk) @ Vb[:, 0], A_one.blk_oracle(blk.to(dev)) @ Vb[:, 0], rtol=2e-5, atol=2e-5))
        ok &= tuple(A_sh.row_oracle(blk).shape) == (b, n) and tuple(A_sh.blk_oracle(blk).shape) == (b, b)
    # replicated-state solvers over the sharded operator; local seeds differ on purpose
    torch.manual_seed(100 + rank)
    A = sharded_kernel_linop(Xd, Xd, KernelConfig(lengthscale=1.0), "rbf", dev)
    full = RBFLinOp(Xd, Xd, KernelConfig(lengthscale=1.0))
    with replicated_rng():
        W, log = LinSys(A, B.to(dev), reg=0.5).solve(
            PCGConfig(device=dev, max_iters=60, rtol=1e-4, precond_config=NystromConfig(rank=80, rho=0.5, sketch="gauss")),
            torch.zeros(n, k, device=dev), callback_freq=1)
        W2, _ = LinSys(A, B.to(dev), reg=0.5, A_row_oracle=A.row_oracle, A_blk_oracle=A.blk_oracle).solve(
            SAPConfig(device=dev, max_iters=20, rtol=1e-4, blk_sz=300, precond_config=NystromConfig(rank=40, rho=0.5),
                      accel_config=SAPAccelConfig(mu=0.5, nu=2.0)), torch.zeros(n, k, device=dev), callback_freq=10)
    ok &= bool((log[max(log)]["metrics"]["internal_metrics"]["rel_res"] <= 1e-4).all())
    K = ko.kernel_matrix(X, X, "rbf", 1.0, dtype=torch.float64)
    ref = torch.linalg.solve(K + 0.5 * torch.eye(n, dtype=torch.float64), B.double())
    ok &= bool(torch.linalg.norm(W.cpu().double() - ref) <= 1e-3 * torch.linalg.norm(ref))
    for T in (W, W2):
        other = T.clone()
        dist.broadcast(other, src=0)
        ok &= bool(torch.equal(T, other))
    # ---- round 2: row-sharded solver state, fused products on the shards, host delivery ----
    from rlaopt_b200.linops import apply_fused
    from rlaopt_b200.solvers._pcg import PCG
    from rlaopt_b200.solvers._pcg_sharded import ShardedPCG
    from rlaopt_b200.utils import SharedPinnedTensor

    k16 = 16
    B16 = torch.randn(n, k16, generator=g)
    B16[:, :5] = (K @ B16[:, :5].double()).float()  # smooth right-hand sides converge first: partial masks
    res = {}
    for sharded in ("1", "0"):
        for mode in ("true", "recurrence"):
            os.environ["RLAOPT_B200_SHARDED_STATE"] = sharded
            system = LinSys(A, B16.to(dev), reg=0.5)
            torch.manual_seed(200 + rank)
            with replicated_rng():
                Ws, logs = system.solve(
                    PCGConfig(device=dev, max_iters=80, rtol=1e-4, precond_config=NystromConfig(rank=80, rho=0.5, sketch="gauss")),
                    torch.zeros(n, k16, device=dev), callback_freq=1, residual=mode)
            ok &= isinstance(system._solver, ShardedPCG if sharded == "1" else PCG)
            ok &= bool((logs[max(logs)]["metrics"]["internal_metrics"]["rel_res"] <= 1e-4).all())
            res[(sharded, mode)] = (Ws, max(logs))
    os.environ.pop("RLAOPT_B200_SHARDED_STATE", None)
    ref16 = torch.linalg.solve(K + 0.5 * torch.eye(n, dtype=torch.float64), B16.double())
    for key, (Ws, its) in res.items():
        ok &= bool(torch.linalg.norm(Ws.cpu().double() - ref16) <= 2e-3 * torch.linalg.norm(ref16))
        ok &= abs(its - res[("0", "true")][1]) <= 2
    # fused product on the row shards: terms on every rank's rows, Gram and norms all-reduced
    C, Lg = torch.randn(n, k, generator=g), torch.randn(n, 2, generator=g)
    Yf, Gf, Sf = apply_fused(A, V.to(dev), alpha=-1.0, addend=C.to(dev), beta=0.25, rhs=B.to(dev), gamma=1.0,
                             gram_with=Lg.to(dev), want_sqnorm=True)
    reff = B.double() - K @ V.double() + 0.25 * C.double()
    ok &= ko.rel_fro_error(Yf, reff) <= 1e-5 and ko.rel_fro_error(Gf, Lg.double().T @ reff) <= 1e-4
    ok &= ko.rel_fro_error(Sf, (reff * reff).sum(0)) <= 1e-4
    # result delivered into host memory mapped by both ranks, each rank its own rows
    shared = SharedPinnedTensor("spmd_gpu_test", (n, k))
    shared.tensor.zero_()
    dist.barrier()
    A.matmat_to_host(V.to(dev), shared.tensor)
    ok &= ko.rel_fro_error(shared.tensor, K @ V.double()) <= 1e-5
    shared.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SPMD_GPU_OK" if int(flag.item()) == 1 else "SPMD_GPU_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
