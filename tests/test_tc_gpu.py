"""tcgen05 tensor-core path (layout = TC) against the fp64 oracle and against the CUDA-core path.

The TC kernel evaluates the L2 kernels in GEMM form with split-precision tensor-core
products (fp16 hi/lo pairs for X.Y^T and for P.V, with exact power-of-two scales) and fp32 accumulation; the bar is
the same 1e-5 relative Frobenius error as the CUDA-core kernel (BASELINE north_star).
"""
import pytest
import torch

from oracle import kernel_oracle as ko

pytestmark = pytest.mark.gpu

TC_KERNELS = ["rbf", "matern12", "matern32", "matern52"]


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
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f32, 500, 10) == LAYOUT_TC  # wide-d variant (K-blocks through smem)
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f32, 3000, 10) == LAYOUT_SIMT  # d > 2048
    assert ops.choose_layout(ops.KERNEL_IDS["rbf"], f64, 128, 64) == LAYOUT_SIMT  # fp64
    assert ops.choose_layout(ops.KERNEL_IDS["laplace"], f32, 32, 16) == LAYOUT_SIMT  # L1 distance
    assert ops.choose_layout(ops.KERNEL_IDS["matern12"], f32, 32, 16) == LAYOUT_TC  # near pairs recomputed exactly


@pytest.mark.parametrize("name", TC_KERNELS)
@pytest.mark.parametrize(
    "n,m,d,k",
    [(1, 1, 1, 1), (2, 3, 3, 2), (127, 129, 33, 17), (129, 300, 16, 64), (300, 257, 128, 70), (200, 1000, 192, 130),
     (1000, 77, 50, 16), (64, 20000, 8, 1), (5000, 5000, 128, 64), (700, 900, 128, 16), (300, 2000, 100, 1),
     (513, 1000, 64, 32), (40000, 300, 16, 2)],
)
def test_tc_against_fp64_oracle(dev, name, n, m, d, k):
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    A1, A2 = _rand((n, d), 1) / d**0.5, _rand((m, d), 2) / d**0.5
    V, W = _rand((m, k), 3), _rand((n, k), 4)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 0.9, 1.7, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 0.9, 1.7, layout=LAYOUT_TC)
    assert got.shape == (n, k)
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


def test_wide_dynamic_range_v(dev):
    """V tiles carry their own power-of-two scale: rows of V spanning 24 orders of magnitude."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, m, d, k = 512, 4096, 64, 32
    A1, A2 = _rand((n, d), 15) / d**0.5, _rand((m, d), 16) / d**0.5
    V = _rand((m, k), 17) * torch.logspace(-12, 12, m).unsqueeze(1)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-5
    # tiny and huge V as a whole (scale is exact, results scale exactly)
    V1 = _rand((m, k), 18)
    base = kernel_matmat(A1.to(dev), A2.to(dev), V1.to(dev), "matern52", 1.0, layout=LAYOUT_TC)
    for e in (-100, -30, 40, 100):
        got = kernel_matmat(A1.to(dev), A2.to(dev), (V1 * 2.0**e).to(dev), "matern52", 1.0, layout=LAYOUT_TC)
        assert torch.equal(got, base * 2.0**e), e


@pytest.mark.parametrize("name", TC_KERNELS)
def test_tiny_kernel_values_keep_relative_accuracy(dev, name):
    """Far-apart clusters: every K_ij is tiny (down to 1e-27 for RBF).  P is scaled per row and
    sub-tile before the fp16 split, so the result keeps its relative accuracy instead of flushing."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, m, d, k = 384, 4096, 64, 32
    A1 = _rand((n, d), 19) / d**0.5
    V = _rand((m, k), 20)
    for shift in (0.5, 1.0, 1.5):
        A2 = _rand((m, d), 21) / d**0.5 + shift
        ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 1.0, dtype=torch.float64)
        assert ref.abs().max() > 0
        got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 1.0, layout=LAYOUT_TC)
        assert ko.rel_fro_error(got, ref) <= 1e-5, (name, shift, float(ref.abs().max()))


def test_mixed_magnitudes_within_a_row(dev):
    """One near-duplicate column (K = 1) among far columns (K ~ 1e-9) in the same 64-column sub-tile,
    with V weighting the far columns up: the small entries must not be lost next to the large one."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, m, d, k = 256, 1024, 16, 4
    A1 = _rand((n, d), 22) / d**0.5
    A2 = _rand((m, d), 23) / d**0.5 + 1.5
    A2[::64] = A1[: m // 64] + 1e-3  # one close point per sub-tile for the first rows
    V = _rand((m, k), 24)
    V[::64] *= 1e-6
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-5


@pytest.mark.parametrize("m", [64, 65, 128, 129, 192, 64 * 17, 64 * 33 + 1, 40000])
def test_odd_and_even_sub_tile_counts(dev, m):
    """The two epilogue warpgroups alternate over the 64-column sub-tiles; cover 1, 2, 3 ... tiles,
    ragged last tiles and the split-column path whose last split holds a single tile."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, d, k = 70, 50, 7
    A1, A2, V = _rand((n, d), 25) / d**0.5, _rand((m, d), 26) / d**0.5, _rand((m, k), 27)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-5, m


def test_k_chunks_and_wide_features(dev):
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    for n, m, d, k in ((300, 2000, 64, 200), (257, 1500, 192, 65), (130, 700, 129, 1), (500, 900, 16, 1000)):
        A1, A2, V = _rand((n, d), 28) / d**0.5, _rand((m, d), 29) / d**0.5, _rand((m, k), 30)
        ref = ko.kernel_matmat_gemm_form(A1, A2, V, "matern32", 1.3, dtype=torch.float64)
        got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "matern32", 1.3, layout=LAYOUT_TC)
        assert ko.rel_fro_error(got, ref) <= 1e-5, (n, m, d, k)


@pytest.mark.parametrize("d,ls", [(3, 0.02), (16, 0.05), (128, 0.2), (40, 1.0)])
def test_matern12_coincident_and_near_points(dev, d, ls):
    """exp(-r) amplifies the GEMM-form distance error near r = 0 (error eps (|x|^2 + |y|^2) / 2r): the diagonal of
    K(X, X), exact duplicates across operands and near-duplicates are recomputed from direct differences.
    Small lengthscales make K ~ I + (few near neighbours), the regime where those entries dominate Y."""
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, k = 1500, 5
    X = _rand((n, d), 31) / d**0.5
    X[100:200] = X[:100]                                   # exact duplicates
    X[200:300] = X[:100] + 1e-5 * _rand((100, d), 32)      # near duplicates
    X[300:400] = X[:100] + 3e-3 * _rand((100, d), 33)      # close points
    V = _rand((n, k), 34)
    ref = ko.kernel_matmat(X, X, V, "matern12", ls, dtype=torch.float64)  # direct differences in fp64
    got = kernel_matmat(X.to(dev), X.to(dev), V.to(dev), "matern12", ls, layout=LAYOUT_TC)
    simt = kernel_matmat(X.to(dev), X.to(dev), V.to(dev), "matern12", ls, layout=LAYOUT_SIMT)
    assert ko.rel_fro_error(got, ref) <= 1e-5, (d, ls, ko.rel_fro_error(got, ref), ko.rel_fro_error(simt, ref))
    # different operands holding the same points (separate packs, separate scales)
    A2 = torch.cat([X[:700] * 1.0, _rand((300, d), 35) / d**0.5 * 3.0])
    W = _rand((A2.shape[0], k), 36)
    ref2 = ko.kernel_matmat(X, A2, W, "matern12", ls, dtype=torch.float64)
    got2 = kernel_matmat(X.to(dev), A2.to(dev), W.to(dev), "matern12", ls, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got2, ref2) <= 1e-5


def test_matern12_diagonal_is_exact(dev):
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, d = 640, 24
    X = (_rand((n, d), 37) * 5.0).to(dev)  # |x|^2 ~ 600: far points, K = I to fp32 precision
    Y = kernel_matmat(X, X, torch.eye(n, device=dev), "matern12", 0.05, layout=LAYOUT_TC)
    assert torch.equal(Y.diagonal(), torch.ones(n, device=dev))
    assert float((Y - torch.eye(n, device=dev)).abs().max()) <= 1e-6


@pytest.mark.parametrize("name", TC_KERNELS)
@pytest.mark.parametrize("n,m,d,k", [(300, 500, 193, 3), (129, 1000, 256, 64), (1000, 70, 500, 10), (257, 4100, 784, 1),
                                     (64, 64, 1000, 16), (2000, 2000, 320, 33), (128, 129, 2048, 2), (300, 900, 400, 130), (200, 3000, 256, 200)])
def test_wide_feature_variant(dev, name, n, m, d, k):
    """d > 192: X no longer fits TMEM; MMA1 accumulates S over 64-feature K-blocks streamed through smem (two column
    tiles per segment).  Odd tile counts leave a phantom second tile in the last segment."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    A1, A2 = _rand((n, d), 41) / d**0.5, _rand((m, d), 42) / d**0.5
    V, W = _rand((m, k), 43), _rand((n, k), 44)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 1.1, 0.5, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 1.1, 0.5, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-5, (name, n, m, d, k, ko.rel_fro_error(got, ref))
    ref_t = ko.kernel_matmat_gemm_form(A2, A1, W, name, 1.1, 0.5, dtype=torch.float64)
    got_t = kernel_matmat(A1.to(dev), A2.to(dev), W.to(dev), name, 1.1, 0.5, transpose=True, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got_t, ref_t) <= 1e-5


# ---- register-contraction mode (k <= 4: P.V on the CUDA cores in the epilogue, no MMA2) ----
@pytest.mark.parametrize("name", TC_KERNELS)
@pytest.mark.parametrize(
    "n,m,d,k",
    [(1, 1, 1, 1), (70, 64, 8, 1), (70, 65, 8, 1), (129, 64 * 3, 16, 2), (300, 64 * 5 + 7, 100, 3), (257, 64 * 33 + 1, 150, 4),
     (40, 50000, 16, 1), (20000, 300, 32, 1), (1500, 1500, 192, 2)],
)
def test_register_contraction_against_fp64_oracle(dev, name, n, m, d, k, monkeypatch):
    """k = 1 ... 4 with ragged tiles, one to many sub-tiles (three / four epilogue warpgroups alternate over them),
    the split-column launch (small n, large m), d up to the resident-X limit; the MMA2 path on the same inputs."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    A1, A2 = _rand((n, d), 41) / d**0.5, _rand((m, d), 42) / d**0.5
    V, W = _rand((m, k), 43), _rand((n, k), 44)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 1.1, 0.7, dtype=torch.float64)
    ref_t = ko.kernel_matmat_gemm_form(A2, A1, W, name, 1.1, 0.7, dtype=torch.float64)
    for mode in ("1", "0"):
        monkeypatch.setenv("RLAOPT_B200_TC_KV", mode)
        got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 1.1, 0.7, layout=LAYOUT_TC)
        assert got.shape == (n, k)
        tol = 2e-6 if (mode == "1" and d <= 128) else 1e-5  # 128 < d <= 192 takes the MMA2 path in either mode
        assert ko.rel_fro_error(got, ref) <= tol, (name, n, m, d, k, mode)
        got_t = kernel_matmat(A1.to(dev), A2.to(dev), W.to(dev), name, 1.1, 0.7, transpose=True, layout=LAYOUT_TC)
        assert ko.rel_fro_error(got_t, ref_t) <= tol


def test_register_contraction_vector_operand_and_gather(dev, monkeypatch):
    """1-D operand with row / column index gathers (the SAP / ASkotch oracles), against the oracle and the MMA2 path."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    n, d = 3000, 16
    X = _rand((n, d), 45) / d**0.5
    v = _rand((n,), 46)
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(47))[:700]
    cols = torch.randperm(n, generator=torch.Generator().manual_seed(48))[:1100]
    for name in TC_KERNELS:
        ref = ko.kernel_matmat_gemm_form(X[rows], X[cols], v[cols, None], name, 1.0, dtype=torch.float64)[:, 0]
        outs = []
        for mode in ("1", "0"):
            monkeypatch.setenv("RLAOPT_B200_TC_KV", mode)
            got = kernel_matmat(X.to(dev), X.to(dev), v[cols].to(dev), name, 1.0, row_idx=rows.to(dev), col_idx=cols.to(dev),
                                layout=LAYOUT_TC)
            assert got.shape == (700,)
            assert ko.rel_fro_error(got, ref) <= (2e-6 if mode == "1" else 1e-5), (name, mode)
            outs.append(got)
        assert ko.rel_fro_error(outs[0], outs[1].double().cpu()) <= 3e-6, name


def test_register_contraction_long_positive_sum_and_tiny_values(dev):
    """Three-level summation over 6000 sub-tiles of positive terms; rows whose kernel values are all ~1e-30 keep
    fp32 relative accuracy (no fp16 range involved on this path)."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    g = torch.Generator().manual_seed(49)
    A1, A2 = _rand((300, 8), 50) / 8**0.5, _rand((384000, 8), 51) / 8**0.5
    V = torch.rand(384000, 1, generator=g)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
    got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=LAYOUT_TC)
    assert ko.rel_fro_error(got, ref) <= 1e-6
    far = A1 + 4.2  # squared distances ~ 140: K ~ 1e-30
    ref = ko.kernel_matmat_gemm_form(far, A2[:5000], V[:5000], "rbf", 1.0, dtype=torch.float64)
    # |x|^2 ~ 140 is beyond the tensor-core accuracy budget (ops.tc_accuracy_ok): with the automatic layout the
    # guard sends these rows to the direct-difference kernel and the 1e-5 bar holds (values ~1e-30)
    got = kernel_matmat(far.to(dev), A2[:5000].to(dev), V[:5000].to(dev), "rbf", 1.0)
    assert ko.rel_fro_error(got, ref) <= 1e-5


@pytest.mark.parametrize("name", ["rbf", "matern32", "matern52", "matern12"])
@pytest.mark.parametrize("n,m,d,k", [(300, 1000, 64, 130), (129, 257, 16, 256), (1000, 5000, 128, 257), (260, 700, 33, 1000),
                                     (513, 129, 100, 300), (40000, 300, 8, 384), (77, 64 * 40 + 3, 64, 129)])
def test_two_chunks_per_sub_tile_family(dev, name, n, m, d, k, monkeypatch):
    """k > 128: one CTA contracts the same P' with two 128-column chunks of V (DUAL instantiations; Matern-1/2 keeps one
    chunk per CTA).  Odd chunk counts, ragged last chunks, single sub-tiles, column splits; against the fp64 oracle and
    against the one-chunk kernels on the same inputs."""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    A1, A2 = _rand((n, d), 61) / d**0.5, _rand((m, d), 62) / d**0.5
    V = _rand((m, k), 63)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, 1.1, 0.9, dtype=torch.float64)
    outs = []
    # 2 = wherever possible (the default uses it for 64 < d <= 128 only), 3 = sliced schedule where three S/P buffers fit
    # (d <= 64; elsewhere the same kernels as 2), 0 = never
    for dual in ("2", "3", "0"):
        monkeypatch.setenv("RLAOPT_B200_TC_DUAL", dual)
        got = kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, 1.1, 0.9, layout=LAYOUT_TC)
        assert got.shape == (n, k)
        assert ko.rel_fro_error(got, ref) <= 1e-5, (name, n, m, d, k, dual)
        outs.append(got)
    assert ko.rel_fro_error(outs[0], outs[2].double().cpu()) <= 2e-6
    assert ko.rel_fro_error(outs[1], outs[2].double().cpu()) <= 2e-6


@pytest.mark.parametrize("name,n,m,d,k", [("rbf", 20000, 200_000, 64, 512), ("matern52", 9000, 150_000, 32, 256),
                                          ("rbf", 12000, 100_000, 128, 384)])
def test_two_chunk_kernels_reproduce_the_one_chunk_kernel_bitwise(dev, name, n, m, d, k, monkeypatch):
    """The two-chunk kernels do the same arithmetic in the same order as the one-chunk kernel, so the outputs must be
    IDENTICAL; repeated launches at a size with hundreds of CTAs and column splits.  (An earlier schedule of the sliced
    mode announced P' while the previous sub-tile's second chunk was still in flight and was wrong in about one CTA in 500
    -- profiles/r02_tc_dual_sliced.md; this is the test that finds it.)"""
    from rlaopt_b200._lib import LAYOUT_TC
    from rlaopt_b200.ops import kernel_matmat

    g = torch.Generator(device=dev).manual_seed(5)
    A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
    V = torch.randn(m, k, generator=g, device=dev)
    A1 = A2[:n]
    monkeypatch.setenv("RLAOPT_B200_TC_DUAL", "0")
    Y0 = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
    for dual in ("1", "2", "3"):
        monkeypatch.setenv("RLAOPT_B200_TC_DUAL", dual)
        for _ in range(3):
            Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
            assert torch.equal(Y, Y0), (name, dual, int((Y != Y0).sum()))
