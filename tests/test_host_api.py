"""Host-side interface parity with rlaopt.linops / rlaopt.kernels (CPU, no GPU needed).

The operator arithmetic is replaced by the ``oracle_backend`` stand-in (conftest);
what is under test is the Python layer: rank dispatch, transposes, scaling,
oracles, validation — the logic of ``rlaopt/linops/base.py:95-111``,
``linops/simple.py``, ``linops/mixins.py`` and ``kernels/base.py:23-128``.  The
kernel part replays ``tests/kernels/test_standard.py:172-326`` of the reference.
"""
import pytest
import torch

from oracle import kernel_oracle as ko
from rlaopt_b200.kernels import (
    KernelConfig,
    LaplaceLinOp,
    Matern12LinOp,
    Matern32LinOp,
    Matern52LinOp,
    RBFLinOp,
)
from rlaopt_b200.linops import LinOp, ScaleMixin, SymmetricLinOp, TwoSidedLinOp
from rlaopt_b200.linops.base import _BaseLinOp

CPU = torch.device("cpu")
KERNELS = [
    (RBFLinOp, "rbf"),
    (LaplaceLinOp, "laplace"),
    (Matern12LinOp, "matern12"),
    (Matern32LinOp, "matern32"),
    (Matern52LinOp, "matern52"),
]
TOL = {torch.float32: dict(rtol=1e-4, atol=1e-4), torch.float64: dict(rtol=1e-8, atol=1e-8)}


# ---------------------------------------------------------------- linops
def _dense_ops(M):
    return dict(matvec=lambda x: M @ x, rmatvec=lambda x: M.T @ x)


def test_linop_rank_dispatch_and_vmap_fallback():
    M = torch.randn(4, 3)
    op = LinOp(CPU, torch.Size((4, 3)), matvec=lambda x: M @ x)
    v, X = torch.randn(3), torch.randn(3, 5)
    assert torch.allclose(op @ v, M @ v)
    assert torch.allclose(op @ X, M @ X, atol=1e-6)  # matmat derived with vmap
    with pytest.raises(ValueError, match="1D or 2D"):
        op @ torch.randn(3, 2, 2)
    with pytest.raises(NotImplementedError):
        op.T
    with pytest.raises(NotImplementedError):
        torch.randn(4) @ op


def test_two_sided_and_symmetric():
    M = torch.randn(4, 3, dtype=torch.float64)
    op = TwoSidedLinOp(CPU, torch.Size((4, 3)), dtype=torch.float64, **_dense_ops(M))
    w, W = torch.randn(4, dtype=torch.float64), torch.randn(2, 4, dtype=torch.float64)
    assert torch.allclose(w @ op, w @ M)
    assert torch.allclose(W @ op, W @ M)
    assert op.T.shape == (3, 4) and op.T.dtype == torch.float64
    assert torch.allclose(op.T @ w, M.T @ w)
    assert torch.allclose(op.T.T @ torch.ones(3, dtype=torch.float64), M.sum(1))
    S = M.T @ M
    sym = SymmetricLinOp(CPU, torch.Size((3, 3)), matvec=lambda x: S @ x, dtype=torch.float64)
    assert sym.T is sym
    assert torch.allclose(torch.ones(3, dtype=torch.float64) @ sym, S.sum(0))
    with pytest.raises(ValueError, match="square"):
        SymmetricLinOp(CPU, torch.Size((3, 4)), matvec=lambda x: x)


def test_base_validation():
    with pytest.raises(TypeError):
        LinOp("cpu", torch.Size((2, 2)), matvec=lambda x: x)
    with pytest.raises(TypeError):
        LinOp(CPU, (2, 2), matvec=lambda x: x)
    with pytest.raises(ValueError, match="two elements"):
        LinOp(CPU, torch.Size((2, 2, 2)), matvec=lambda x: x)
    with pytest.raises(ValueError, match="positive"):
        LinOp(CPU, torch.Size((0, 2)), matvec=lambda x: x)
    with pytest.raises(ValueError, match="float32 or torch.float64"):
        LinOp(CPU, torch.Size((2, 2)), matvec=lambda x: x, dtype=torch.float16)
    with pytest.raises(TypeError, match="callable"):
        LinOp(CPU, torch.Size((2, 2)), matvec=3)


def test_scale_mixin():
    class S(ScaleMixin):
        pass

    s = S()
    s._initialize_scaling(2.0)
    f = s._apply_scaling(lambda x: x + 1)
    assert f(1.0) == 4.0
    g = s._apply_scaling(f)  # scales compose
    assert g(1.0) == 8.0
    assert s._apply_scaling(torch.ones(2)).tolist() == [2.0, 2.0]
    s._initialize_scaling(1.0)
    fn = lambda x: x  # noqa: E731
    assert s._apply_scaling(fn) is fn
    s._initialize_scaling(None)
    assert s._scaling == 1.0


# ---------------------------------------------------------------- KernelConfig
def test_kernel_config_validation_and_to():
    cfg = KernelConfig(lengthscale=1.5)
    assert cfg.const_scaling == 1.0 and cfg.to(CPU) is cfg
    assert cfg.to_dict() == {"const_scaling": 1.0, "lengthscale": 1.5}
    with pytest.raises(TypeError):
        KernelConfig(const_scaling=2, lengthscale=1.0)  # ints are rejected like the reference
    with pytest.raises(TypeError):
        KernelConfig(lengthscale=1)
    with pytest.raises(ValueError, match="1 dimension"):
        KernelConfig(lengthscale=torch.ones(2, 2))
    with pytest.raises(TypeError):
        KernelConfig(1.0, 2.0)  # keyword-only
    t = KernelConfig(const_scaling=3.0, lengthscale=torch.tensor([1.0, 2.0]))
    t2 = t.to(CPU)
    assert t2 is not t and torch.equal(t2.lengthscale, t.lengthscale) and t2.const_scaling == 3.0


# ---------------------------------------------------------------- kernel operators
@pytest.fixture(params=[torch.float32, torch.float64], ids=["float32", "float64"])
def precision(request):
    return request.param


@pytest.fixture(params=["scalar", "tensor"])
def kernel_config(request, precision):
    if request.param == "scalar":
        return KernelConfig(const_scaling=2.0, lengthscale=1.0)
    return KernelConfig(const_scaling=2.0, lengthscale=torch.tensor([1.0, 2.0, 3.0], dtype=precision))


@pytest.fixture
def mats(precision):
    g = torch.Generator().manual_seed(1)
    return torch.randn(10, 3, generator=g).to(precision), torch.randn(5, 3, generator=g).to(precision)


@pytest.mark.parametrize("cls,name", KERNELS, ids=[k[1] for k in KERNELS])
def test_kernel_initialisation_and_checks(cls, name, mats, kernel_config):
    A1, A2 = mats
    K = cls(A1, A2, kernel_config=kernel_config)
    assert K.A1.shape == A1.shape and K.A2.shape == A2.shape
    assert K.kernel_config == kernel_config and K.dtype == A1.dtype
    assert K.shape == (10, 5) and K.device == A1.device and K._scaling == 2.0
    assert type(K).__name__.lower() == f"{name}linop"
    with pytest.raises(TypeError):
        cls([[1.0]], A2, kernel_config)
    with pytest.raises(ValueError, match="2D"):
        cls(A1[0], A2, kernel_config)
    with pytest.raises(ValueError, match="same dtype"):
        cls(A1, A2.to(torch.float64 if A1.dtype == torch.float32 else torch.float32), kernel_config)
    with pytest.raises(TypeError, match="KernelConfig"):
        cls(A1, A2, {"lengthscale": 1.0})


def test_cpu_tensors_have_no_fallback(mats):
    A1, A2 = mats
    K = RBFLinOp(A1, A2, KernelConfig(lengthscale=1.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K @ torch.randn(5, dtype=A1.dtype)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        torch.ops.rlaopt_b200.kernel_matmat(A1, A2, torch.randn(5, dtype=A1.dtype), 0, 1.0, None, 1.0)


@pytest.mark.parametrize("cls,name", KERNELS, ids=[k[1] for k in KERNELS])
def test_matmul_transpose_and_oracles(cls, name, mats, kernel_config, oracle_backend):
    """Replay of the reference's TestKernelLinOps on the host layer."""
    A1, A2 = mats
    tol = TOL[A1.dtype]
    K = cls(A1, A2, kernel_config=kernel_config)
    Kd = ko.kernel_matrix(A1, A2, name, kernel_config.lengthscale, kernel_config.const_scaling)
    g = torch.Generator().manual_seed(2)
    v = torch.randn(5, generator=g).to(A1.dtype)
    M = torch.randn(5, 2, generator=g).to(A1.dtype)
    w = torch.randn(10, generator=g).to(A1.dtype)
    W = torch.randn(2, 10, generator=g).to(A1.dtype)
    assert torch.allclose(K @ v, Kd @ v, **tol)
    assert torch.allclose(K @ M, Kd @ M, **tol)
    assert torch.allclose(w @ K, w @ Kd, **tol)
    assert torch.allclose(K.T @ w, Kd.T @ w, **tol)
    assert torch.allclose(W @ K, W @ Kd, **tol)
    assert torch.allclose(K.T @ W.T, Kd.T @ W.T, **tol)
    assert K.T.dtype == K.dtype
    # the two packs are made once and reused by every product above
    assert oracle_backend["pack"] == 2 and oracle_backend["matmat"] == 6

    blk = torch.tensor([0, 1], dtype=torch.long)
    row = K.row_oracle(blk)
    assert isinstance(row, _BaseLinOp) and row.shape == (2, 5)
    assert row.device == K.device and row.dtype == K.dtype
    assert torch.allclose(row @ v, Kd[blk] @ v, **tol)
    assert torch.allclose(row @ M, Kd[blk] @ M, **tol)
    sub = K.blk_oracle(blk)
    assert isinstance(sub, _BaseLinOp) and sub.shape == (2, 2)
    Kb = ko.kernel_matrix(A1[blk], A2[blk], name, kernel_config.lengthscale, kernel_config.const_scaling)
    assert torch.allclose(sub @ v[:2], Kb @ v[:2], **tol)
    assert torch.allclose(sub @ M[:2], Kb @ M[:2], **tol)


def test_oracle_packs_are_memoised_per_block(mats, oracle_backend):
    A1, _ = mats
    K = RBFLinOp(A1, A1, KernelConfig(lengthscale=1.0))
    blk = torch.tensor([3, 1, 4], dtype=torch.long)
    v = torch.randn(3, dtype=A1.dtype)
    for _ in range(4):  # SAP rebuilds the oracle on every power iteration (sap.py:96-97)
        K.blk_oracle(blk) @ v
    assert oracle_backend["pack"] == 1 and oracle_backend["matmat"] == 4
    K.blk_oracle(torch.tensor([3, 1, 4], dtype=torch.long)) @ v  # a different tensor object re-packs
    assert oracle_backend["pack"] == 2


def test_shared_operand_is_packed_once(mats, oracle_backend):
    A1, _ = mats
    K = Matern52LinOp(A1, A1, KernelConfig(lengthscale=2.0))
    K @ torch.randn(10, dtype=A1.dtype)
    torch.randn(10, dtype=A1.dtype) @ K
    assert oracle_backend["pack"] == 1


def test_gather_index_validation_is_host_logic():
    """``A1[blk]`` semantics for the oracles' index lists: range errors raise IndexError before any
    kernel is launched; negative entries are legal (they wrap in the pack kernels)."""
    import pytest
    import torch

    from rlaopt_b200 import ops

    cpu = torch.device("cpu")
    out = ops._check_index(torch.tensor([0, -1, 4, -5]), 5, cpu)
    assert out.dtype == torch.int64 and out.tolist() == [0, -1, 4, -5]
    assert ops._check_index(torch.tensor([], dtype=torch.int64), 5, cpu).numel() == 0
    for bad in ([5], [-6], [0, 1, 99]):
        with pytest.raises(IndexError, match="out of bounds"):
            ops._check_index(torch.tensor(bad), 5, cpu)
    with pytest.raises(ValueError):
        ops._check_index(torch.zeros(2, 2, dtype=torch.int64), 5, cpu)
    with pytest.raises(IndexError):
        ops._check_index(torch.tensor([0.0, 1.0]), 5, cpu)


def test_pack_cache_tracks_in_place_edits(oracle_backend):
    """The pack cache is keyed on the in-place versions of A1, A2 and the lengthscale tensor."""
    import torch

    from rlaopt_b200.kernels import KernelConfig, RBFLinOp

    X = torch.randn(20, 3)
    ls = torch.ones(3)
    op = RBFLinOp(X, X, KernelConfig(lengthscale=ls))
    v = torch.randn(20)
    op @ v
    op @ v
    assert oracle_backend["pack"] == 1  # one shared pack, reused
    X.mul_(2.0)
    op @ v
    assert oracle_backend["pack"] == 2  # data edited in place: repacked
    ls.add_(1.0)
    op @ v
    assert oracle_backend["pack"] == 3  # lengthscale edited in place: repacked
    blk = torch.tensor([1, 2, 3])
    ro = op.row_oracle(blk)
    ro @ v
    ro @ v
    packs = oracle_backend["pack"]
    X.add_(1.0)
    ro @ v
    assert oracle_backend["pack"] == packs + 2  # block pack and column pack both rebuilt


def test_tc_accuracy_budget():
    from rlaopt_b200 import ops

    assert ops.tc_accuracy_ok(ops.KERNEL_IDS["rbf"], 30.0, 30.0)
    assert not ops.tc_accuracy_ok(ops.KERNEL_IDS["rbf"], 40.0, 40.0)
    assert ops.tc_norm_budget(ops.KERNEL_IDS["matern32"]) < ops.tc_norm_budget(ops.KERNEL_IDS["matern52"]) \
        < ops.tc_norm_budget(ops.KERNEL_IDS["rbf"])
