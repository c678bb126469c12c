"""The preconditioner / sketch API checks of test_preconditioners_api.py on cuda:0, plus Nystrom built on the fused
kernel operator (the sketch Y = K @ Omega is one kernel matmat with k = rank)."""
import pytest
import torch

import test_preconditioners_api as api

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _on_gpu(monkeypatch):
    monkeypatch.setattr(api, "DEVICE", torch.device("cuda:0"))


@pytest.fixture(params=[torch.float32, torch.float64], ids=["float32", "float64"])
def precision(request):
    return request.param


@pytest.fixture
def spd(precision):
    torch.manual_seed(0)
    A = torch.randn(50, 50, device="cuda:0", dtype=precision)
    return A @ A.T


@pytest.mark.parametrize("sketch", ["gauss", "ortho"])
@pytest.mark.parametrize("as_linop", [False, True], ids=["tensor", "linop"])
def test_nystrom(spd, precision, sketch, as_linop):
    api.test_nystrom(spd, precision, sketch, as_linop)


@pytest.mark.parametrize("as_linop", [False, True], ids=["tensor", "linop"])
def test_newton(spd, precision, as_linop):
    api.test_newton(spd, precision, as_linop)


@pytest.mark.parametrize("name", ["gauss", "ortho"])
def test_sketches(name, precision):
    api.test_sketch_shapes_and_application(name, precision)


@pytest.mark.parametrize("rank", [40, 200])
def test_nystrom_on_kernel_operator(rank):
    """Nystrom of an RBF kernel operator: U S U^T matches the Nystrom approximation computed densely in fp64 from
    the same Omega, and P^-1 (K + rho I) has a small condition number on the captured subspace."""
    from oracle import kernel_oracle as ko
    from rlaopt_b200.kernels import KernelConfig, RBFLinOp
    from rlaopt_b200.preconditioners import NystromConfig
    from rlaopt_b200.preconditioners.nystrom import Nystrom
    from rlaopt_b200.utils import host_rng

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n, d = 3000, 6
    X = torch.randn(n, d, generator=g) / d**0.5
    A = RBFLinOp(X.to(dev), X.to(dev), KernelConfig(lengthscale=1.0))
    torch.manual_seed(5)
    with host_rng():
        P = Nystrom(NystromConfig(rank=rank, rho=1e-2, sketch="gauss", damping_mode="non_adaptive"))
        P._update(A, dev)
    torch.manual_seed(5)
    Omega = (torch.randn(rank, n) / rank**0.5).T.double()  # the same draw (gauss.py:46-52)
    K = ko.kernel_matrix(X, X, "rbf", 1.0, dtype=torch.float64)
    Y = K @ Omega
    core = Omega.T @ Y
    nys = Y @ torch.linalg.solve(core + 1e-12 * torch.eye(rank, dtype=torch.float64), Y.T)
    got = ((P.U * P.S) @ P.U.T).cpu().double()
    assert torch.linalg.norm(got - nys) <= 2e-3 * torch.linalg.norm(nys)  # fp32 build incl. the eps*trace shift
    # fp32 round trip: the error grows with S_max / rho (~ n / rho for a kernel matrix), so check it at rho = 1
    P.config.rho, P.L = 1.0, None
    v = torch.randn(n, 4, generator=g).to(dev)
    back = P @ (P._inv @ v)
    assert torch.linalg.norm(back - v) <= 5e-3 * torch.linalg.norm(v)
