"""tcgen05 tensor-core path (layout = TC) against the fp64 oracle and against the CUDA-core path.

The TC kernel evaluates the L2 kernels in GEMM form with split-precision tensor-core
products (fp16 hi/lo for X.Y^T, tf32 hi/lo for P.V) and fp32 accumulation; the bar is
the same 1e-5 relative Frobenius error as the CUDA-core kernel (BASELINE north_star).
"""
import pytest
import torch

from oracle import kernel_oracle as ko

pytestmark = pytest.mark.gpu

TC_KERNELS = ["rbf", "matern32", "matern52"]


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_layout_selection(dev):
    from rlaopt_b200 import ops
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC

    f32, f64 = torch.float32, torch.float64
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f32, 128, 64) == LAYOUT_TC
    assert ops.choose_layout(ops.KERNEL_IDS["matern52"], f32, 3, 1) == LAYOUT_TC
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f32, 500, 10) == LAYOUT_SIMT  # d > 192
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f64, 128, 64) == LAYOUT_SIMT  # fp64
    assert ops.choose_layout(ops.KERNEL_IDS["laplace"], f32, 32, 16) == LAYOUT_SIMT  # L1 distance
    assert ops.choose_layout(ops.KERNEL_IDS["matern12"], f32, 32, 16) == LAYOUT_SIMT  # non-smooth at r = 0


@pytest.mark.parametrize("name", TC_KERNELS + ["matern12"])
@pytest.mark.parametrize(
    "n,m,d,k",
    [(1, 1, 1, 1), (2, 3, 3, 2), (127, 129, 33, 17), (129, 300, 16, 64), (300, 257, 128, 70), (200, 1000, 192, 130),
     (1000, 77, 50, 16), (64, 20000, 8, 1), (5000, 5000, 128, 64)],
)
def test_tc_against_fp64_oracle(dev, name, n, m, d, k):
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    A1, A2 = _rand((n, d), 1) / d**0.5, _rand((m, d), 2) / d**0.5
    V, W = _rand((m, k), 3), _rand((n, k), 4)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 0.9, 1.7, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 0.9, 1.7, layout=LAYOUT_TC)
    assert got.shape == (n, k)
    # Matern-1/2 is only offered on this path on request; away from r = 0 it meets the same bar
    assert ko.rel_fro_error(got, ref) <= 1e-5, (name, n, m, d, k)
    ref_t = ko.kernel_matmat_gemm_form(A2, A1, W, name, 0.9, 1.7, dtype=torch.float64)
    got_t = kernel_matmat(A1.to(dev), A2.to(dev), W.to(dev), name, 0.9, 1.7, transpose=True, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got_t, ref_t) <= 1e-5


def test_tc_matches_cuda_core_path(dev):
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, d, k = 4096, 64, 32
    X = (_rand((n, d), 5) / d**0.5).to(dev)
    V = _rand((n, k), 6).to(dev)
    ls = torch.linspace(0.7, 1.5, d, device=dev)  # per-feature lengthscale
    for name in TC_KERNELS:
        a = kernel_matmat(X, X, V, name, ls, 2.0, layout=LAYOUT_TC)
        b = kernel_matmat(X, X, V, name, ls, 2.0, layout=LAYOUT_SIMT)
        assert ko.rel_fro_error(a, b) <= 2e-6, name


def test_tc_scale_invariance_and_offsets(dev):
    """Per-operand power-of-two scaling: huge / tiny feature magnitudes with a matching lengthscale,
    and a common offset (uncentred data) that the GEMM-form distance must survive."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, m, d, k = 600, 900, 40, 8
    A1, A2, V = _rand((n, d), 7) / d**0.5, _rand((m, d), 8) / d**0.5, _rand((m, k), 9)
    base = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    for s in (1e-12, 1e-3, 1e4, 1e15):
        got = kernel_matmat((A1 * s).to(dev), (A2 * s).to(dev), V.to(dev), "rbf", float(s), layout=LAYOUT_TC)
        assert ko.rel_fro_error(got, base) <= 1e-5, s
    shift = 3.0 / d**0.5  # |x|^2 grows from ~1 to ~10
    ref = ko.kernel_matmat_gemm_form(A1 + shift, A2 + shift, V, "matern52", 1.0, dtype=torch.float64)
    got = kernel_matmat((A1 + shift).to(dev), (A2 + shift).to(dev), V.to(dev), "matern52", 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-5


def test_tc_long_sums_do_not_drift(dev):
    """All-positive V over 200k columns: the tensor core's truncating accumulator would bias the
    sum by ~1e-5; per-sub-tile fp32 drains keep the mean relative error below 1e-6."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, m, d, k = 512, 200_000, 32, 16
    A1, A2 = _rand((n, d), 10) / d**0.5, _rand((m, d), 11) / d**0.5
    V = _rand((m, k), 12).abs()
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=LAYOUT_TC).cpu().double()
    rel = (got - ref) / ref
    assert rel.abs().max().item() <= 5e-6
    assert abs(rel.mean().item()) <= 1e-6


def test_tc_deterministic(dev):
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    X = (_rand((3000, 24), 13) / 5).to(dev)
    V = _rand((3000, 5), 14).to(dev)
    a = kernel_matmat(X, X, V, "matern32", 1.0, layout=LAYOUT_TC)
    b = kernel_matmat(X, X, V, "matern32", 1.0, layout=LAYOUT_TC)
    assert torch.equal(a, b)
