"""Host logic of the distributed operators on CPU "devices" (no GPU needed).

Replays ``tests/kernels/test_distributed.py:117-423`` of the reference (RBF, scalar
lengthscale, one CPU worker, full-kernel and oracle-only modes) on the in-process
multi-device implementation, with the shard arithmetic supplied by the oracle
stand-in.  Also covers the generic ``DistributedLinOp`` protocol
(``rlaopt/linops/distributed.py:40-50,82-115``) with several shards.
"""
import pytest
import torch

from oracle import kernel_oracle as ko
from rlaopt_b200.kernels import DistributedMatern32LinOp, DistributedRBFLinOp, KernelConfig
from rlaopt_b200.linops import (
    DistributedLinOp,
    DistributedSymmetricLinOp,
    DistributedTwoSidedLinOp,
    LinOp,
    SymmetricLinOp,
    TwoSidedLinOp,
)
from rlaopt_b200.linops.base import _BaseLinOp

CPU = torch.device("cpu")
TOL = dict(rtol=1e-4, atol=1e-4)


def _dense_shard(M):
    return TwoSidedLinOp(CPU, torch.Size(M.shape), lambda x: M @ x, lambda x: M.T @ x, lambda x: M @ x, lambda x: M.T @ x)


@pytest.mark.parametrize("mode", ["row", "column"])
def test_generic_distributed_protocol(mode):
    g = torch.Generator().manual_seed(0)
    M = torch.randn(11, 7, generator=g)
    pieces = torch.chunk(M, 3, dim=0 if mode == "row" else 1)
    op = DistributedTwoSidedLinOp(torch.Size(M.shape), [_dense_shard(p) for p in pieces], mode)
    v, V = torch.randn(7, generator=g), torch.randn(7, 4, generator=g)
    w, W = torch.randn(11, generator=g), torch.randn(4, 11, generator=g)
    assert torch.allclose(op @ v, M @ v, atol=1e-5)
    assert torch.allclose(op @ V, M @ V, atol=1e-5)
    assert torch.allclose(w @ op, w @ M, atol=1e-5)
    assert torch.allclose(W @ op, W @ M, atol=1e-5)
    assert torch.allclose(op.T @ w, M.T @ w, atol=1e-5)
    assert torch.allclose(op.T.T @ v, M @ v, atol=1e-5)
    assert op.devices == [CPU] * 3
    with pytest.raises(AttributeError, match="devices"):
        op.device
    op.shutdown()
    op.shutdown()  # idempotent
    with pytest.raises(RuntimeError, match="shut down"):
        op @ v


def test_distributed_validation():
    shard = LinOp(CPU, torch.Size((2, 2)), matvec=lambda x: x)
    with pytest.raises(TypeError):
        DistributedLinOp(torch.Size((2, 2)), (shard,), "row")
    with pytest.raises(ValueError, match="_BaseLinOp"):
        DistributedLinOp(torch.Size((2, 2)), [object()], "row")
    with pytest.raises(ValueError, match="Invalid value"):
        DistributedLinOp(torch.Size((2, 2)), [shard], "diagonal")
    with pytest.raises(ValueError, match="same dtype"):
        DistributedLinOp(
            torch.Size((4, 2)),
            [shard, LinOp(CPU, torch.Size((2, 2)), matvec=lambda x: x, dtype=torch.float64)],
            "row",
        )
    S = torch.eye(2)
    sym = SymmetricLinOp(CPU, torch.Size((2, 2)), matvec=lambda x: S @ x)
    dsym = DistributedSymmetricLinOp(torch.Size((2, 2)), [sym], "row")
    assert dsym.T is dsym
    with pytest.raises(ValueError, match="square"):
        DistributedSymmetricLinOp(torch.Size((2, 3)), [sym], "row")


@pytest.fixture
def data():
    g = torch.Generator().manual_seed(3)
    return torch.randn(10, 3, generator=g), torch.randn(5, 3, generator=g)


@pytest.fixture
def cfg():
    return KernelConfig(const_scaling=2.0, lengthscale=1.0)


def test_distributed_kernel_attributes_and_matmul(data, cfg, oracle_backend):
    A1, A2 = data
    K = DistributedRBFLinOp(A1, A2, kernel_config=cfg, devices={CPU})
    try:
        assert K._scaling == 2.0 and K.shape == (10, 5) and K.dtype == torch.float32
        assert len(K.A1_row_chunks) == 1 and torch.equal(K.A1_row_chunks[0], torch.arange(10))
        assert len(K.A2_row_chunks) == 1 and len(K.A2_chunks) == 1 and len(K.kernel_ops) == 1
        assert K.devices == [CPU] and K.A1 is A1 and K.A2 is A2 and K.kernel_config is cfg
        with pytest.raises(AttributeError):
            K.device
        Kd = ko.kernel_matrix(A1, A2, "rbf", 1.0, 2.0)
        v, M = torch.randn(5), torch.randn(5, 2)
        w, W = torch.randn(10), torch.randn(2, 10)
        assert torch.allclose(K @ v, Kd @ v, **TOL)
        assert torch.allclose(K @ M, Kd @ M, **TOL)
        assert torch.allclose(w @ K, w @ Kd, **TOL)
        assert torch.allclose(K.T @ w, Kd.T @ w, **TOL)
        assert torch.allclose(W @ K, W @ Kd, **TOL)
        assert torch.allclose(K.T @ W.T, Kd.T @ W.T, **TOL)
    finally:
        K.shutdown()


@pytest.mark.parametrize("use_full_kernel", [True, False])
def test_distributed_oracles(data, cfg, use_full_kernel, oracle_backend):
    A1, A2 = data
    K = DistributedRBFLinOp(A1, A2, kernel_config=cfg, devices={CPU}, use_full_kernel=use_full_kernel)
    try:
        blk = torch.tensor([0, 1], dtype=torch.long)
        v, M = torch.randn(5), torch.randn(5, 2)
        row = K.row_oracle(blk)
        assert isinstance(row, _BaseLinOp) and row.shape == (2, 5) and row.dtype == K.dtype
        Kr = ko.kernel_matrix(A1[blk], A2, "rbf", 1.0, 2.0)
        assert torch.allclose(row @ v, Kr @ v, **TOL)
        assert torch.allclose(row @ M, Kr @ M, **TOL)
        sub = K.blk_oracle(blk)
        assert isinstance(sub, _BaseLinOp) and sub.shape == (2, 2)
        Kb = ko.kernel_matrix(A1[blk], A2[blk], "rbf", 1.0, 2.0)
        assert torch.allclose(sub @ v[:2], Kb @ v[:2], **TOL)
        assert torch.allclose(sub @ M[:2], Kb @ M[:2], **TOL)
        if not use_full_kernel:
            assert all(type(op) is TwoSidedLinOp for op in K.kernel_ops)  # shape-only placeholders
            with pytest.raises(RuntimeError, match="use_full_kernel=False"):
                K @ v
    finally:
        K.shutdown()


def test_distributed_kernel_validation(data, cfg):
    A1, A2 = data
    with pytest.raises(TypeError, match="set"):
        DistributedRBFLinOp(A1, A2, cfg, devices=[CPU])
    with pytest.raises(ValueError, match="non-empty"):
        DistributedRBFLinOp(A1, A2, cfg, devices=set())
    with pytest.raises(ValueError, match="torch.device"):
        DistributedRBFLinOp(A1, A2, cfg, devices={"cpu"})
    with pytest.raises(ValueError, match="2D"):
        DistributedRBFLinOp(A1[0], A2, cfg, devices={CPU})


def test_row_partition_matches_reference_chunking(cfg, oracle_backend, monkeypatch):
    """Several 'devices': partition = torch.chunk(arange(n), g) (kernels/base.py:297-302,462)."""
    # three distinct CPU device objects stand in for three GPUs
    devs = [torch.device("cpu"), torch.device("cpu", 0), torch.device("meta")]
    devs = devs[:2]  # cpu and cpu:0 compare unequal as set members but both execute on the host
    g = torch.Generator().manual_seed(5)
    A = torch.randn(11, 4, generator=g)
    K = DistributedMatern32LinOp(A, A, KernelConfig(lengthscale=1.3), devices=set(devs))
    try:
        expect = torch.chunk(torch.arange(11), len(devs))
        assert [c.tolist() for c in K.A1_row_chunks] == [c.tolist() for c in expect]
        assert [op.shape[0] for op in K.kernel_ops] == [len(c) for c in expect]
        Kd = ko.kernel_matrix(A, A, "matern32", 1.3)
        V = torch.randn(11, 3, generator=g)
        assert torch.allclose(K @ V, Kd @ V, **TOL)
        assert torch.allclose(K.T @ V, Kd @ V, **TOL)
        blk = torch.tensor([10, 2, 5, 7, 0], dtype=torch.long)
        assert torch.allclose(K.row_oracle(blk) @ V, Kd[blk] @ V, **TOL)
        assert torch.allclose(K.blk_oracle(blk) @ V[:5], Kd[blk][:, blk] @ V[:5], **TOL)
        # fewer rows than devices: chunk() yields fewer shards, extra devices stay idle
        tiny = DistributedRBFLinOp(A[:1], A, KernelConfig(lengthscale=1.0), devices=set(devs))
        assert len(tiny.kernel_ops) == 1 and torch.allclose(tiny @ V, ko.kernel_matrix(A[:1], A, "rbf", 1.0) @ V, **TOL)
    finally:
        K.shutdown()
