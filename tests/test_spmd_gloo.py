"""World-size-2 (and 3) gloo runs of the SPMD row-sharded operator on CPU.

Covers the N>1 plumbing bench.py uses under torchrun — row partition, padded
all-gather of ragged row blocks, all-reduce of transpose partials — with the shard
arithmetic supplied by the oracle (the CUDA kernel cannot run here).
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, m, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kernel_oracle as ko
        from rlaopt_b200.linops import TwoSidedLinOp
        from rlaopt_b200.linops.spmd import RowShardedLinOp, shard_rows

        g = torch.Generator().manual_seed(0)  # same data on every rank
        A1, A2 = torch.randn(n, 4, generator=g), torch.randn(m, 4, generator=g)
        V, W = torch.randn(m, 3, generator=g), torch.randn(n, 3, generator=g)
        Kd = ko.kernel_matrix(A1, A2, "matern52", 1.2, 1.5)
        lo, hi = shard_rows(n, world)[rank]
        assert [list(c) for c in torch.chunk(torch.arange(n), world)] == [
            list(range(a, b)) for a, b in shard_rows(n, world) if b > a
        ]
        local = None
        if hi > lo:
            Kr = Kd[lo:hi]
            local = TwoSidedLinOp(torch.device("cpu"), torch.Size(Kr.shape), lambda x: Kr @ x, lambda x: Kr.T @ x,
                                  lambda x: Kr @ x, lambda x: Kr.T @ x)
        op = RowShardedLinOp(local, torch.Size((n, m)), torch.device("cpu"), torch.float32)
        ok = True
        ok &= torch.allclose(op @ V, Kd @ V, atol=1e-5)
        ok &= torch.allclose(op @ V[:, 0], Kd @ V[:, 0], atol=1e-5)
        ok &= torch.allclose(op.T @ W, Kd.T @ W, atol=1e-5)
        ok &= torch.allclose(W.T @ op, W.T @ Kd, atol=1e-5)
        ok &= torch.allclose(op.T.T @ V, Kd @ V, atol=1e-5)
        ok &= op.local_matmat(V).shape[0] == hi - lo
        # block-CG style global dot products: allreduce of per-rank partial Grams
        P = (op @ V)[lo:hi]
        gram = V.new_zeros(3, 3) if hi == lo else W[lo:hi].T @ P
        dist.all_reduce(gram)
        ok &= torch.allclose(gram, W.T @ (Kd @ V), atol=1e-4)
        # host -> every rank replication (1/world of the rows per rank, then all-gather), ragged n included
        from rlaopt_b200.kernels.sharded import replicate_from_host

        ok &= torch.equal(replicate_from_host(A1, torch.device("cpu")), A1)
        ok &= torch.equal(replicate_from_host(V[:, 0].contiguous(), torch.device("cpu")), V[:, 0])
        # fused product on the row shards: element-wise terms on every rank's rows, Gram and column norms
        # all-reduced in one buffer (the only cross-rank reductions of a block-PCG step)
        from rlaopt_b200.linops import apply_fused
        from rlaopt_b200.utils import SharedPinnedTensor

        C, B = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
        Y, G, S = apply_fused(op, V, alpha=-1.0, addend=C, beta=0.25, rhs=B, gamma=1.0, gram_with=W, want_sqnorm=True)
        ref = B - Kd @ V + 0.25 * C
        ok &= torch.allclose(Y, ref, atol=1e-5) and torch.allclose(G, W.T @ ref, atol=1e-4)
        ok &= torch.allclose(S, (ref * ref).sum(0), atol=1e-4)
        idx = torch.randperm(n, generator=g)
        Y2, _, S2 = apply_fused(op, V, addend=C, beta=2.0, addend_idx=idx, want_sqnorm=True, store=False)
        ok &= Y2 is None and torch.allclose(S2, ((Kd @ V + 2.0 * C[idx]) ** 2).sum(0), rtol=1e-4)
        # result delivered to host memory mapped by every rank: each rank writes its own row block
        shared = SharedPinnedTensor(f"gloo_test_{n}_{m}", (n, 3))
        shared.tensor.zero_()
        dist.barrier()
        op.matmat_to_host(V, shared.tensor)
        ok &= torch.allclose(shared.tensor, Kd @ V, atol=1e-5)
        shared.close()
        ok &= not os.path.exists(shared.path) or rank != 0
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,m", [(2, 11, 7), (2, 8, 8), (3, 4, 5), (3, 2, 6)])
def test_row_sharded_operator_gloo(world, n, m):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, m, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}


def _solver_worker(rank, world, port, out):
    """KRR solves with replicated solver state over the row-sharded operator: different local seeds on the two
    ranks, random draws taken from rank 0 (replicated_rng) -> identical iterates, equal to a one-process solve."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kernel_oracle as ko
        from rlaopt_b200.linops import LinOp, SymmetricLinOp, TwoSidedLinOp
        from rlaopt_b200.linops.spmd import RowShardedLinOp, shard_rows
        from rlaopt_b200.models import LinSys
        from rlaopt_b200.preconditioners import NystromConfig
        from rlaopt_b200.solvers import PCGConfig, SAPAccelConfig, SAPConfig
        from rlaopt_b200.utils import replicated_rng

        cpu = torch.device("cpu")
        g = torch.Generator().manual_seed(0)
        n, k = 400, 2
        X = torch.randn(n, 5, generator=g, dtype=torch.float64) / 5**0.5
        B = torch.randn(n, k, generator=g, dtype=torch.float64)
        K = ko.kernel_matrix(X, X, "rbf", 1.0, dtype=torch.float64)
        lo, hi = shard_rows(n, world)[rank]
        Kr = K[lo:hi]
        local = TwoSidedLinOp(cpu, torch.Size(Kr.shape), lambda x: Kr @ x, lambda x: Kr.T @ x, lambda x: Kr @ x,
                              lambda x: Kr.T @ x, dtype=torch.float64)
        A = RowShardedLinOp(local, torch.Size((n, n)), cpu, torch.float64)

        def row_oracle(blk):  # sharded over the columns would need a reduce; rows of K are cheap here
            Kb = K[blk]
            return LinOp(cpu, torch.Size((len(blk), n)), lambda v: Kb @ v, lambda V: Kb @ V, dtype=torch.float64)

        def blk_oracle(blk):
            Kbb = K[blk][:, blk]
            return LinOp(cpu, torch.Size((len(blk), len(blk))), lambda v: Kbb @ v, lambda V: Kbb @ V, dtype=torch.float64)

        torch.manual_seed(1234 + rank)  # deliberately different local streams
        ok = True
        with replicated_rng():
            W, log = LinSys(A, B, reg=0.3).solve(
                PCGConfig(device=cpu, max_iters=60, rtol=1e-10, precond_config=NystromConfig(rank=40, rho=0.3, sketch="gauss")),
                torch.zeros(n, k, dtype=torch.float64), callback_freq=1)
            W2, _ = LinSys(A, B, reg=0.3, A_row_oracle=row_oracle, A_blk_oracle=blk_oracle).solve(
                SAPConfig(device=cpu, max_iters=30, rtol=1e-10, blk_sz=50, precond_config=NystromConfig(rank=20, rho=0.3),
                          accel_config=SAPAccelConfig(mu=0.3, nu=3.0)), torch.zeros(n, k, dtype=torch.float64), callback_freq=10)
        ref = torch.linalg.solve(K + 0.3 * torch.eye(n, dtype=torch.float64), B)
        ok &= bool(torch.linalg.norm(W - ref) <= 1e-8 * torch.linalg.norm(ref))
        # identical iterates on every rank
        for T in (W, W2):
            mine = T.clone()
            other = T.clone()
            dist.broadcast(other, src=0)
            ok &= bool(torch.equal(mine, other))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_spmd_replicated_solvers_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_solver_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _sharded_pcg_worker(rank, world, port, out):
    """Row-sharded solver state (ShardedPCG) against the replicated PCG over the same row-sharded operator: same
    iteration counts, iterates equal up to the summation order of the reductions, columns converging at different
    iterations (partial masks), true and recurrence residuals, ragged row blocks (world = 3)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rlaopt_b200.linops import TwoSidedLinOp
        from rlaopt_b200.linops.spmd import RowShardedLinOp, shard_rows
        from rlaopt_b200.models import LinSys
        from rlaopt_b200.preconditioners import IdentityConfig, NystromConfig
        from rlaopt_b200.solvers import PCGConfig
        from rlaopt_b200.solvers._pcg import PCG
        from rlaopt_b200.solvers._pcg_sharded import ShardedPCG
        from rlaopt_b200.utils import replicated_rng

        cpu = torch.device("cpu")
        g = torch.Generator().manual_seed(0)
        n, k = 401, 5
        X = torch.randn(n, 6, generator=g, dtype=torch.float64) / 6**0.5
        sq = (X * X).sum(1)
        K = torch.exp(-0.5 * (sq[:, None] + sq[None, :] - 2 * X @ X.T).clamp_min(0))
        B = torch.randn(n, k, generator=g, dtype=torch.float64)
        B[:, :2] = K @ B[:, :2]  # smooth right-hand sides converge earlier: partial masks
        lo, hi = shard_rows(n, world)[rank]
        Kr = K[lo:hi]
        local = TwoSidedLinOp(cpu, torch.Size(Kr.shape), lambda x: Kr @ x, lambda x: Kr.T @ x, lambda x: Kr @ x,
                              lambda x: Kr.T @ x, dtype=torch.float64)
        A = RowShardedLinOp(local, torch.Size((n, n)), cpu, torch.float64)
        ok = True
        for pc in (lambda: NystromConfig(rank=40, rho=0.3, sketch="gauss"), lambda: NystromConfig(rank=30, rho=0.3),
                   lambda: IdentityConfig()):
            for mode in ("true", "recurrence"):
                res = {}
                for sharded in ("1", "0"):
                    os.environ["RLAOPT_B200_SHARDED_STATE"] = sharded
                    system = LinSys(A, B, reg=0.3)
                    torch.manual_seed(7)
                    with replicated_rng():
                        W, log = system.solve(PCGConfig(device=cpu, max_iters=120, rtol=1e-9, precond_config=pc()),
                                              torch.zeros(n, k, dtype=torch.float64), callback_freq=1, residual=mode)
                    kind = ShardedPCG if sharded == "1" else PCG
                    ok &= isinstance(system._solver, kind)
                    res[sharded] = (W, max(log), torch.stack([log[i]["metrics"]["internal_metrics"]["rel_res"] for i in sorted(log)]))
                # block CG amplifies the rounding difference of the reductions near its floor: the runs may stop
                # one logging period apart, the early residual history is the same
                # block CG amplifies the rounding difference of the reductions (a 1e-16 difference in the summation
                # order of three partials is O(1) ten iterations later without a preconditioner):
                # the first iterations agree, the runs may stop a few logging periods apart, the solutions agree
                ok &= abs(res["1"][1] - res["0"][1]) <= 3
                ok &= bool(torch.allclose(res["1"][2][:4], res["0"][2][:4], rtol=1e-6, atol=0))
                ok &= bool(torch.linalg.norm(res["1"][0] - res["0"][0]) <= 1e-8 * torch.linalg.norm(res["0"][0]))
                ref = torch.linalg.solve(K + 0.3 * torch.eye(n, dtype=torch.float64), B)
                ok &= bool(torch.linalg.norm(res["1"][0] - ref) <= 1e-7 * torch.linalg.norm(ref))
        os.environ.pop("RLAOPT_B200_SHARDED_STATE", None)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_state_pcg_matches_replicated_pcg(world):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sharded_pcg_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}
