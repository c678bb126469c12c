"""Fused output stage of the kernel matmat: ``alpha c K V + beta C[ci] + gamma B[bi]``,
the Gram matrix ``L^T Y`` and the squared column norms in one pass, through the C entry
``rlaopt_b200_matmat_packed_fused_*`` -- against the same quantities assembled from the fp64 oracle."""
import pytest
import torch

from oracle import kernel_oracle as ko

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name,dtype", [("rbf", torch.float32), ("matern52", torch.float32), ("laplace", torch.float32),
                                        ("rbf", torch.float64), ("matern32", torch.float64)])
@pytest.mark.parametrize("n,m,d,k", [(300, 500, 9, 1), (1000, 40000, 16, 3), (777, 900, 32, 16), (2100, 1300, 64, 64),
                                     (20000, 700, 8, 5)])
def test_fused_terms_gram_and_norms(dev, name, dtype, n, m, d, k):
    from rlaopt_b200 import ops

    tol = 1e-5 if dtype == torch.float32 else 1e-11
    A1, A2 = (_rand((n, d), 1) / d**0.5).to(dtype), (_rand((m, d), 2) / d**0.5).to(dtype)
    V, C, B, L = (_rand(s, i).to(dtype) for i, s in enumerate([(m, k), (n, k), (n, k), (n, min(k, 7))], start=3))
    kid = ops.KERNEL_IDS[name]
    layout = ops.choose_layout(kid, dtype, d, k)
    c = ops.column_mean(A2.to(dev)) if layout == ops.LAYOUT_TC else None
    P1, P2 = ops.pack_points(A1.to(dev), 1.2, None, layout, c), ops.pack_points(A2.to(dev), 1.2, None, layout, c)
    KV = ko.kernel_matmat(A1, A2, V, name, 1.2, 1.5, dtype=torch.float64)
    ref = -0.7 * KV + 0.3 * C.double() - 1.1 * B.double()
    Y, G, S = ops.matmat_packed_fused(P1, P2, V.to(dev), name, 1.5, alpha=-0.7, addend=C.to(dev), beta=0.3,
                                      rhs=B.to(dev), gamma=-1.1, gram_with=L.to(dev), want_sqnorm=True)
    assert Y.shape == (n, k) and G.shape == (L.shape[1], k) and S.shape == (k,)
    assert ko.rel_fro_error(Y, ref) <= tol
    assert ko.rel_fro_error(G, L.double().T @ ref) <= 10 * tol  # a sum over n rows of products of O(1) terms
    assert ko.rel_fro_error(S, (ref * ref).sum(0)) <= 10 * tol
    # reductions only: nothing n x k is written
    Y2, G2, S2 = ops.matmat_packed_fused(P1, P2, V.to(dev), name, 1.5, alpha=-0.7, addend=C.to(dev), beta=0.3,
                                         rhs=B.to(dev), gamma=-1.1, gram_with=L.to(dev), want_sqnorm=True, store=False)
    assert Y2 is None and torch.equal(G2, G) and torch.equal(S2, S)  # deterministic reductions
    # plain product through the fused entry == the ordinary entry
    Y3, _, _ = ops.matmat_packed_fused(P1, P2, V.to(dev), name, 1.5)
    assert ko.rel_fro_error(Y3, ops.matmat_packed(P1, P2, V.to(dev), name, 1.5).double().cpu()) <= 1e-6


def test_fused_gathered_terms_block_gradient(dev):
    """``K[blk, :] W + reg W[blk] - B[blk]`` (``rlaopt/solvers/sap.py:113-127``) from the row oracle in one pass,
    with negative indices and a vector operand."""
    from rlaopt_b200.kernels import KernelConfig, Matern52LinOp, RBFLinOp
    from rlaopt_b200.linops import apply_fused

    n, d = 5000, 12
    X = _rand((n, d), 11) / d**0.5
    for k in (1, 4, 20):
        W, B = _rand((n, k), 12), _rand((n, k), 13)
        blk = torch.randperm(n, generator=torch.Generator().manual_seed(14))[:600]
        blk[:3] -= n  # negative indices wrap
        for cls, name in ((RBFLinOp, "rbf"), (Matern52LinOp, "matern52")):
            op = cls(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0, const_scaling=0.8))
            got, _, _ = apply_fused(op.row_oracle(blk), W.to(dev), addend=W.to(dev), beta=0.05, addend_idx=blk,
                                    rhs=B.to(dev), gamma=-1.0, rhs_idx=blk)
            ref = ko.kernel_matmat(X[blk], X, W, name, 1.0, 0.8, dtype=torch.float64) + 0.05 * W[blk].double() - B[blk].double()
            assert ko.rel_fro_error(got, ref) <= 1e-5, (name, k)
    w, b = _rand((n,), 15), _rand((n,), 16)
    op = RBFLinOp(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0))
    got, _, sq = apply_fused(op, w.to(dev), alpha=-1.0, addend=w.to(dev), beta=-0.1, rhs=b.to(dev), gamma=1.0, want_sqnorm=True)
    ref = b.double() - (ko.kernel_matmat(X, X, w[:, None], "rbf", 1.0, dtype=torch.float64)[:, 0] + 0.1 * w.double())
    assert got.shape == (n,) and ko.rel_fro_error(got, ref) <= 1e-5
    assert abs(float(sq[0]) - float((ref * ref).sum())) <= 1e-5 * float((ref * ref).sum())


def test_fused_wide_k_falls_back_to_separate_reductions(dev):
    """k > 64: the element-wise terms stay fused, Gram and norms are taken by separate passes (same numbers)."""
    from rlaopt_b200.kernels import KernelConfig, RBFLinOp
    from rlaopt_b200.linops import apply_fused

    n, d, k = 1500, 16, 100
    X, V = _rand((n, d), 21) / d**0.5, _rand((n, k), 22)
    op = RBFLinOp(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0))
    Y, G, S = apply_fused(op, V.to(dev), addend=V.to(dev), beta=0.5, gram_with=V.to(dev), want_sqnorm=True)
    ref = ko.kernel_matmat(X, X, V, "rbf", 1.0, dtype=torch.float64) + 0.5 * V.double()
    assert ko.rel_fro_error(Y, ref) <= 1e-5
    assert ko.rel_fro_error(G, V.double().T @ ref) <= 1e-4 and ko.rel_fro_error(S, (ref * ref).sum(0)) <= 1e-4


@pytest.mark.parametrize("k", [1, 3])
def test_recurrence_residual_halves_the_products_per_logged_iteration(dev, k):
    """Block PCG at ``callback_freq = 1``: with ``residual="recurrence"`` a logged iteration costs one kernel product
    instead of two (``rlaopt/models/linsys.py:96-99``); same iteration count, same solution, and the stop is
    confirmed by one true residual."""
    from rlaopt_b200 import ops
    from rlaopt_b200.kernels import KernelConfig, RBFLinOp
    from rlaopt_b200.models import LinSys
    from rlaopt_b200.preconditioners import NystromConfig
    from rlaopt_b200.solvers import PCGConfig
    from rlaopt_b200.utils import host_rng

    n, d, reg = 6000, 8, 0.5
    X, B = (_rand((n, d), 31) / d**0.5).to(dev), _rand((n, k), 32).to(dev)
    runs = {}
    for mode in ("true", "recurrence"):
        calls = {"n": 0}
        orig = ops.matmat_packed_fused

        def counting(*a, _orig=orig, _calls=calls, **kw):
            _calls["n"] += 1
            return _orig(*a, **kw)

        ops.matmat_packed_fused = counting
        try:
            A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
            system = LinSys(A, B, reg=reg)
            cfg = PCGConfig(device=dev, max_iters=60, rtol=1e-4, precond_config=NystromConfig(rank=100, rho=reg, sketch="gauss"))
            torch.manual_seed(3)
            with host_rng():
                before = calls["n"]
                W, log = system.solve(cfg, torch.zeros(n, k, device=dev), callback_freq=1, residual=mode)
        finally:
            ops.matmat_packed_fused = orig
        iters = max(log)
        runs[mode] = (W, iters, calls["n"] - before, float(log[iters]["metrics"]["internal_metrics"]["rel_res"].max()))
    (W_t, it_t, n_t, r_t), (W_r, it_r, n_r, r_r) = runs["true"], runs["recurrence"]
    assert it_t == it_r
    assert r_t <= 1e-4 and r_r <= 1e-4  # the recurrence run's last metric is the confirming true residual
    assert ko.rel_fro_error(W_r, W_t.double().cpu()) <= 1e-4
    # products: sketch (1, plain entry, not counted) + initial residual (1) + per iteration 1 (+1 metric in "true" mode)
    assert n_t == 1 + 1 + 2 * it_t       # init residual, metric at 0, then step + metric per iteration
    assert n_r == 1 + it_r + 1           # init residual, one product per step, one confirming residual


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_edge_shapes(dev, dtype):
    """Tiny and ragged shapes, the widest reductions (k = gram_cols = 64: 64 KB of shared memory in fp64), an empty
    column set (the product is an empty sum: only the addend / rhs terms remain), single rows."""
    from rlaopt_b200 import ops

    tol = 1e-5 if dtype == torch.float32 else 1e-11
    for n, m, d, k, gc in ((1, 1, 1, 1, 1), (3, 130, 5, 2, 2), (65, 7, 3, 64, 64), (200, 300, 40, 64, 64), (129, 0, 4, 3, 2)):
        A1, A2 = _rand((n, d), 1).to(dtype), _rand((max(m, 1), d), 2).to(dtype)[:m]
        V, C, B, L = (_rand(s, i).to(dtype) for i, s in enumerate([(max(m, 1), k), (n, k), (n, k), (n, gc)], start=3))
        V = V[:m]
        layout = ops.choose_layout(0, dtype, d, k)
        c = ops.column_mean(A2.to(dev)) if (layout == ops.LAYOUT_TC and m > 0) else None
        P1 = ops.pack_points(A1.to(dev), 1.0, None, layout, c)
        P2 = ops.pack_points(A2.to(dev), 1.0, None, layout, c)
        Y, G, S = ops.matmat_packed_fused(P1, P2, V.to(dev), "rbf", 2.0, alpha=0.5, addend=C.to(dev), beta=-1.5,
                                          rhs=B.to(dev), gamma=0.25, gram_with=L.to(dev), want_sqnorm=True)
        KV = ko.kernel_matmat(A1, A2, V, "rbf", 1.0, 2.0, dtype=torch.float64) if m > 0 else torch.zeros(n, k, dtype=torch.float64)
        ref = 0.5 * KV - 1.5 * C.double() + 0.25 * B.double()
        assert Y.shape == (n, k) and G.shape == (gc, k) and S.shape == (k,)
        assert ko.rel_fro_error(Y, ref) <= tol, (n, m, d, k)
        assert ko.rel_fro_error(G, L.double().T @ ref) <= 10 * tol and ko.rel_fro_error(S, (ref * ref).sum(0)) <= 10 * tol
    # beyond the reduction range the C entry refuses instead of computing something else
    P = ops.pack_points(_rand((10, 3), 9).to(dev), 1.0, None, ops.LAYOUT_SIMT)
    with pytest.raises(RuntimeError, match="k <= 64"):
        ops.matmat_packed_fused(P, P, _rand((10, 70), 10).to(dev), "rbf", want_sqnorm=True)
    with pytest.raises(ValueError, match="rows"):
        ops.matmat_packed_fused(P, P, _rand((10, 2), 10).to(dev), "rbf", addend=_rand((9, 2), 11).to(dev), beta=1.0)
