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
