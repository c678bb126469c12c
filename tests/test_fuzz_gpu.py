"""Seeded random sweep of the fused kernel matmat through the C ABI against the fp64 oracle: random kernel,
shapes (ragged tiles, one to many row blocks, k from 1 to a few k-chunks, d up to beyond the tensor-core limit),
scalar / per-feature lengthscale, const_scaling, transpose, row / column index gathers, 1-D right-hand sides and
strided operands.  Exercises every instantiation family of the tensor-core kernel (2 / 3 epilogue warpgroups,
64- / 128-column chunks, Matern-1/2 variant), the split-column path and the CUDA-core kernel."""
import random

import pytest
import torch

from oracle import kernel_oracle as ko

pytestmark = pytest.mark.gpu

KERNELS = ["rbf", "laplace", "matern12", "matern32", "matern52"]


def _case(seed):
    rng = random.Random(seed)
    kernel = rng.choice(KERNELS)
    n = rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 255, 257, 300, 511, 700, 1025, 2500])
    m = rng.choice([1, 3, 63, 64, 65, 127, 128, 129, 130, 200, 1000, 1100, 4097, 9000])
    d = rng.choice([1, 2, 3, 7, 8, 15, 16, 17, 31, 32, 33, 50, 63, 64, 65, 100, 127, 128, 129, 191, 192, 193, 260])
    k = rng.choice([1, 1, 2, 3, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 128, 129, 200, 257])
    return dict(kernel=kernel, n=n, m=m, d=d, k=k, ls_vec=rng.random() < 0.3, scale=rng.choice([1.0, 1.0, 2.5, -0.7]),
                transpose=rng.random() < 0.3, rows=rng.random() < 0.25, cols=rng.random() < 0.25,
                vec=rng.random() < 0.15, strided=rng.random() < 0.2, same=rng.random() < 0.3)


@pytest.mark.parametrize("seed", range(120))
def test_random_case_against_fp64_oracle(seed):
    from rlaopt_b200.ops import kernel_matmat

    c = _case(seed)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1000 + seed)
    n, m, d, k = c["n"], c["m"], c["d"], c["k"]
    if c["vec"]:
        k = 1
    A1 = torch.randn(n, d, generator=g) / d**0.5
    A2 = A1 if (c["same"] and not c["rows"] and not c["cols"]) else torch.randn(m, d, generator=g) / d**0.5
    m = A2.shape[0]
    ls = (0.5 + torch.rand(d, generator=g)) if c["ls_vec"] else 0.6 + 0.8 * float(torch.rand(1, generator=g))
    row_idx = torch.randint(0, n, (max(1, n // 2),), generator=g) if c["rows"] else None
    col_idx = torch.randint(0, m, (max(1, (2 * m) // 3),), generator=g) if c["cols"] else None
    n_eff = n if row_idx is None else row_idx.numel()
    m_eff = m if col_idx is None else col_idx.numel()
    rows_in = n_eff if c["transpose"] else m_eff
    V = torch.randn(rows_in, k, generator=g)
    if c["strided"]:  # non-contiguous right-hand side: a column slice of a wider matrix
        wide = torch.randn(rows_in, k + 3, generator=g)
        Vd = wide.to(dev)[:, 1:1 + k]
        V = wide[:, 1:1 + k]
    else:
        Vd = V.to(dev)
    if c["vec"]:
        V, Vd = V[:, 0], Vd[:, 0]
    A1e = A1 if row_idx is None else A1[row_idx]
    A2e = A2 if col_idx is None else A2[col_idx]
    if c["transpose"]:
        ref = ko.kernel_matmat(A2e, A1e, V, c["kernel"], ls, c["scale"], dtype=torch.float64)
    else:
        ref = ko.kernel_matmat(A1e, A2e, V, c["kernel"], ls, c["scale"], dtype=torch.float64)
    lsd = ls.to(dev) if isinstance(ls, torch.Tensor) else ls
    A1d = A1.to(dev)
    A2d = A1d if A2 is A1 else A2.to(dev)
    got = kernel_matmat(A1d, A2d, Vd, c["kernel"], lsd, c["scale"], transpose=c["transpose"],
                        row_idx=None if row_idx is None else row_idx.to(dev),
                        col_idx=None if col_idx is None else col_idx.to(dev))
    assert got.shape == ref.shape, (c, got.shape, ref.shape)
    err = ko.rel_fro_error(got, ref)
    assert err <= 1e-5, (c, err)


@pytest.mark.parametrize("seed", range(200, 240))
def test_random_case_with_cta_pairs_and_column_chunks(seed, monkeypatch):
    """Same sweep with the launch-level options forced on small problems: CTA pairs (cluster multicast, phantom
    row block when the count is odd) and 16-tile column chunks (many partial results, single-tile last chunk)."""
    monkeypatch.setenv("RLAOPT_B200_TC_PAIR", "1")
    monkeypatch.setenv("RLAOPT_B200_TC_SPLIT_TILES", "16")
    test_random_case_against_fp64_oracle(seed)
