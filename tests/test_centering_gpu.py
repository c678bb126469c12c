"""Uncentred / large-norm data on the tensor-core path (VERDICT r1 "what's weak" #1).

The GEMM-form distance |x|^2 + |y|^2 - 2 x.y carries an absolute error ~3e-7 (|x|^2 + |y|^2).  All five kernels
are functions of x - y (``rlaopt/kernels/standard.py:31-43``), so the tensor-core packs subtract one common
center (the column means of A2) from both operands; what is left is guarded at run time: when the centred norms
still exceed the budget of ``ops.tc_accuracy_ok`` the operator runs on the direct-difference CUDA-core kernel.
Bar everywhere: 1e-5 relative Frobenius error against the fp64 oracle (BASELINE north_star).
"""
import pytest
import torch

from oracle import kernel_oracle as ko

pytestmark = pytest.mark.gpu

TC_KERNELS = ["rbf", "matern12", "matern32", "matern52"]
TOL = 1e-5  # north_star: fp32 matmat within 1e-5 (relative Frobenius) of the reference formulas in fp64


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _shifted(n, m, d, shift, seed):
    """Standardised features with sigma = 1/sqrt(d) and a common mean of `shift` sigma per feature."""
    off = (_rand((d,), seed + 7).sign() * shift) / d**0.5
    A1 = _rand((n, d), seed) / d**0.5 + off
    A2 = _rand((m, d), seed + 1) / d**0.5 + off
    return A1, A2


@pytest.mark.parametrize("name", TC_KERNELS)
@pytest.mark.parametrize("shift", [5.0, 10.0, 100.0])
@pytest.mark.parametrize("d", [8, 32, 128])
@pytest.mark.parametrize("k", [1, 64])
def test_mean_shifted_data_through_the_operator(dev, name, shift, d, k):
    """A1 != A2, mean 5 / 10 / 100 sigma: forward, transpose, row and block oracles stay within 1e-5 and stay on
    the tensor-core layout (the shift is removed by the common center, not by falling back)."""
    from rlaopt_b200 import kernels, ops
    from rlaopt_b200._lib import LAYOUT_TC

    n, m = 700, 1500
    A1, A2 = _shifted(n, m, d, shift, 100 + d)
    V, W = _rand((m, k), 3), _rand((n, k), 4)
    cls = getattr(kernels, {"rbf": "RBFLinOp", "matern12": "Matern12LinOp", "matern32": "Matern32LinOp",
                            "matern52": "Matern52LinOp"}[name])
    op = cls(A1.to(dev), A2.to(dev), kernels.KernelConfig(lengthscale=1.0, const_scaling=1.3))
    assert op._layout_for(V.to(dev)) == LAYOUT_TC
    got = op @ V.to(dev)
    ref = ko.kernel_matmat(A1, A2, V, name, 1.0, 1.3, dtype=torch.float64)
    assert ko.rel_fro_error(got, ref) <= TOL, ("forward", ko.rel_fro_error(got, ref))
    got_t = op.T @ W.to(dev)
    ref_t = ko.kernel_matmat(A2, A1, W, name, 1.0, 1.3, dtype=torch.float64)
    assert ko.rel_fro_error(got_t, ref_t) <= TOL, ("transpose", ko.rel_fro_error(got_t, ref_t))
    blk = torch.randperm(n, generator=torch.Generator().manual_seed(5))[:300]
    got_r = op.row_oracle(blk) @ V.to(dev)
    assert ko.rel_fro_error(got_r, ref[blk]) <= TOL, ("row oracle", ko.rel_fro_error(got_r, ref[blk]))
    Vb = _rand((300, k), 6)
    got_b = op.blk_oracle(blk) @ Vb.to(dev)
    ref_b = ko.kernel_matmat(A1[blk], A2[blk], Vb, name, 1.0, 1.3, dtype=torch.float64)
    assert ko.rel_fro_error(got_b, ref_b) <= TOL, ("blk oracle", ko.rel_fro_error(got_b, ref_b))


@pytest.mark.parametrize("name", TC_KERNELS)
def test_mean_shifted_symmetric_operator_one_shot_entry(dev, name):
    """K(X, X) with a 10-sigma mean through the one-shot C entry (centres on the device, in its workspace)."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, d, k = 3000, 32, 16
    X, _ = _shifted(n, 1, d, 10.0, 11)
    V = _rand((n, k), 12)
    ref = ko.kernel_matmat(X, X, V, name, 1.0, dtype=torch.float64)
    got = kernel_matmat(X.to(dev), X.to(dev), V.to(dev), name, 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= TOL
    # per-feature lengthscale and a gathered block, shifted data
    ls = torch.linspace(0.8, 1.6, d)
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(13))[:500]
    ref = ko.kernel_matmat(X[rows], X, V, name, ls, dtype=torch.float64)
    got = kernel_matmat(X.to(dev), X.to(dev), V.to(dev), name, ls.to(dev), row_idx=rows.to(dev), layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= TOL


def test_c_abi_center_argument_and_statistics(dev):
    """``rlaopt_b200_column_mean_f32`` / the ``center`` argument of ``rlaopt_b200_pack_points_f32`` /
    ``rlaopt_b200_packed_stats_host`` called through ctypes."""
    from rlaopt_b200 import _lib, ops
    from rlaopt_b200._lib import LAYOUT_TC

    n, d = 5000, 24
    X = (_rand((n, d), 21) / d**0.5 + 3.0).to(dev)
    c = ops.column_mean(X)
    assert torch.allclose(c.cpu().double(), X.cpu().double().mean(0), rtol=0, atol=1e-6)
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(22))[:777].to(dev)
    ci = ops.column_mean(X, idx)
    assert torch.allclose(ci.cpu().double(), X[idx].cpu().double().mean(0), rtol=0, atol=1e-6)
    assert torch.equal(ops.column_mean(X), c)  # fixed summation order: bit-reproducible
    P = ops.pack_points(X, 1.0, None, LAYOUT_TC, c)
    want = float(((X.double() - c.double()) ** 2).sum(1).max())
    assert abs(P.max_sqnorm - want) <= 1e-5 * want
    P0 = ops.pack_points(X, 1.0, None, LAYOUT_TC)  # uncentred: the statistic shows the offset
    assert P0.max_sqnorm > 9.0 * d * 0.9
    assert P.stats()[1] == 0
    lib = _lib.load()
    assert lib.rlaopt_b200_abi_version() == 2


def test_guard_falls_back_to_direct_differences(dev):
    """Centred norms beyond the tensor-core budget (|x / l|^2 ~ 130 here): the operator leaves the tensor-core
    layout on its own and keeps the parity bar; the diagonal of K(X, X) is exactly c (direct differences)."""
    from rlaopt_b200 import ops
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC
    from rlaopt_b200.kernels import KernelConfig, Matern32LinOp, RBFLinOp

    n, d, k = 1200, 128, 8
    X = _rand((n, d), 31)  # unit-variance features with lengthscale 1: |x|^2 ~ d
    V = _rand((n, k), 32)
    for cls, name in ((RBFLinOp, "rbf"), (Matern32LinOp, "matern32")):
        op = cls(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0))
        assert ops.choose_layout(ops.KERNEL_IDS[name], torch.float32, d, k) == LAYOUT_TC
        assert op._layout_for(V.to(dev)) == LAYOUT_SIMT
        got = op @ V.to(dev)
        ref = ko.kernel_matmat(X, X, V, name, 1.0, dtype=torch.float64)
        assert ko.rel_fro_error(got, ref) <= TOL
        blk = torch.arange(0, n, 3)
        got_b = op.blk_oracle(blk) @ V[blk].to(dev)
        ref_b = ko.kernel_matmat(X[blk], X[blk], V[blk], name, 1.0, dtype=torch.float64)
        assert ko.rel_fro_error(got_b, ref_b) <= TOL
    # with a lengthscale that normalises the data the same operator is back on the tensor cores
    op = RBFLinOp(X.to(dev), X.to(dev), KernelConfig(lengthscale=float(d) ** 0.5))
    assert op._layout_for(V.to(dev)) == LAYOUT_TC
    ref = ko.kernel_matmat(X, X, V, "rbf", float(d) ** 0.5, dtype=torch.float64)
    assert ko.rel_fro_error(op @ V.to(dev), ref) <= TOL


def test_guard_budget_is_calibrated(dev):
    """The budget of ``ops.tc_norm_budget`` is not loose: forcing the tensor-core layout right at the budget stays
    within the bar for every kernel (error measured, norms measured from the pack header)."""
    from rlaopt_b200 import ops
    from rlaopt_b200._lib import LAYOUT_TC

    n, d, k = 2000, 32, 16
    Z = _rand((n, d), 41) / d**0.5
    V = _rand((n, k), 42)
    for name in TC_KERNELS:
        kid = ops.KERNEL_IDS[name]
        budget = ops.tc_norm_budget(kid)
        # scale the cloud so that 2 max|x|^2 sits just under the budget
        s = (0.5 * budget / float((Z.double() ** 2).sum(1).max())) ** 0.5 * 0.98
        X = (Z * s).float()
        c = torch.zeros(d)
        P = ops.pack_points(X.to(dev), 1.0, None, LAYOUT_TC, c.to(dev))
        assert ops.tc_accuracy_ok(kid, P.max_sqnorm, P.max_sqnorm)
        got = ops.matmat_packed(P, P, V.to(dev), name)
        ref = ko.kernel_matmat(X, X, V, name, 1.0, dtype=torch.float64)
        assert ko.rel_fro_error(got, ref) <= TOL, (name, ko.rel_fro_error(got, ref))


def test_packs_follow_in_place_updates(dev):
    """The reference's LazyTensor reads the live tensors; the pack cache is keyed on their in-place versions."""
    from rlaopt_b200.kernels import KernelConfig, RBFLinOp

    n, d, k = 600, 16, 4
    X = (_rand((n, d), 51) / d**0.5).to(dev)
    V = _rand((n, k), 52).to(dev)
    ls = torch.full((d,), 1.0, device=dev)
    cfg = KernelConfig(lengthscale=ls)
    op = RBFLinOp(X, X, cfg)
    y0 = op @ V
    X.mul_(1.5)  # in-place edit of the data
    y1 = op @ V
    ref1 = ko.kernel_matmat(X.cpu(), X.cpu(), V.cpu(), "rbf", ls.cpu(), dtype=torch.float64)
    assert ko.rel_fro_error(y1, ref1) <= TOL
    assert ko.rel_fro_error(y0, ref1) > 1e-3
    ls.mul_(2.0)  # in-place edit of the lengthscale tensor
    y2 = op @ V
    ref2 = ko.kernel_matmat(X.cpu(), X.cpu(), V.cpu(), "rbf", ls.cpu(), dtype=torch.float64)
    assert ko.rel_fro_error(y2, ref2) <= TOL
    blk = torch.arange(0, n, 2)
    r0 = op.row_oracle(blk) @ V
    X.add_(0.25)
    r1 = op.row_oracle(blk) @ V  # same blk object: the memo must notice the new data version
    ref3 = ko.kernel_matmat(X.cpu()[blk], X.cpu(), V.cpu(), "rbf", ls.cpu(), dtype=torch.float64)
    assert ko.rel_fro_error(r1, ref3) <= TOL
    assert r0.shape == r1.shape


def test_gather_indices_are_validated(dev):
    """``A1[blk]`` semantics: negative indices wrap, out-of-range host indices raise IndexError, and the
    pack kernels never read out of bounds (a bad device index packs as a zero point and is counted)."""
    from rlaopt_b200 import ops
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC
    from rlaopt_b200.kernels import KernelConfig, LaplaceLinOp, RBFLinOp

    n, d, k = 300, 8, 3
    X = _rand((n, d), 61) / d**0.5
    V = _rand((n, k), 62)
    for cls, name in ((RBFLinOp, "rbf"), (LaplaceLinOp, "laplace")):
        op = cls(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0))
        blk = torch.tensor([0, -1, 5, -300, 299])
        ref = ko.kernel_matmat(X[blk], X, V, name, 1.0, dtype=torch.float64)
        assert ko.rel_fro_error(op.row_oracle(blk) @ V.to(dev), ref) <= TOL
        for bad in (torch.tensor([0, 300]), torch.tensor([-301, 2])):
            with pytest.raises(IndexError):
                op.row_oracle(bad) @ V.to(dev)
            with pytest.raises(IndexError):
                op.blk_oracle(bad) @ V[:2].to(dev)
    # C ABI level: an out-of-range index is memory-safe and reported by the pack statistics
    lib = ops._lib.load()
    Xd = X.to(dev)
    idx = torch.tensor([1, 10_000_000, -10_000_000, 2], device=dev)
    for layout in (LAYOUT_TC, LAYOUT_SIMT):
        nbytes = lib.rlaopt_b200_packed_bytes(4, d, 4, layout)
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        rc = lib.rlaopt_b200_pack_points_f32(Xd.data_ptr(), 4, n, d, d, idx.data_ptr(), 1.0, None, None, layout,
                                             buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        if layout == LAYOUT_TC:
            P = ops.PackedPoints(buf, 4, d, torch.float32, layout)
            assert P.stats()[1] == 2


def test_far_field_rows_auto_layout(dev):
    """Two clusters 4.2 apart per feature (kernel values ~1e-30): the centred norms exceed the budget, the one-shot
    entry with automatic layout evaluates them by direct differences."""
    from rlaopt_b200.ops import kernel_matmat

    g = torch.Generator().manual_seed(49)
    A1, A2 = _rand((300, 8), 50) / 8**0.5, _rand((5000, 8), 51) / 8**0.5
    V = torch.rand(5000, 1, generator=g)
    far = A1 + 4.2
    ref = ko.kernel_matmat(far, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(far.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0)
    # fp32 direct differences: D ~ 140 carries eps D ~ 1e-5 absolute, i.e. ~5e-6 relative in exp(-D/2)
    assert ko.rel_fro_error(got, ref) <= TOL
