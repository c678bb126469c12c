"""API behaviour of the sketches and preconditioners, following the reference's own test plan
(tests/preconditioners/test_{identity,newton,nystrom,preconditioner}.py): shapes / dtypes / devices,
orthonormal U, non-negative S, products against the explicit formulas, P @ (P^-1 x) = x, tensors and linear
operators as input, fixed vs adaptive damping, input validation.  Same tolerances as the reference
(1e-4 fp32, 1e-8 fp64).  Runs on the CPU here; test_preconditioners_gpu repeats it on cuda:0.
"""
import pytest
import torch

from rlaopt_b200.linops import SymmetricLinOp
from rlaopt_b200.preconditioners import IdentityConfig, NewtonConfig, NystromConfig, Preconditioner
from rlaopt_b200.preconditioners.identity import Identity
from rlaopt_b200.preconditioners.newton import Newton
from rlaopt_b200.preconditioners.nystrom import Nystrom
from rlaopt_b200.sketches import get_sketch

TOL = {torch.float32: dict(rtol=1e-4, atol=1e-4), torch.float64: dict(rtol=1e-8, atol=1e-8)}
DEVICE = torch.device("cpu")


@pytest.fixture(params=[torch.float32, torch.float64], ids=["float32", "float64"])
def precision(request):
    return request.param


@pytest.fixture
def spd(precision):
    torch.manual_seed(0)
    A = torch.randn(50, 50, device=DEVICE, dtype=precision)
    return A @ A.T


def _as_linop(M):
    return SymmetricLinOp(M.device, M.shape, lambda x: M @ x, dtype=M.dtype)


@pytest.mark.parametrize("sketch", ["gauss", "ortho"])
@pytest.mark.parametrize("as_linop", [False, True], ids=["tensor", "linop"])
def test_nystrom(spd, precision, sketch, as_linop):
    tol = TOL[precision]
    cfg = NystromConfig(rank=20, sketch=sketch, rho=1e0, damping_mode="non_adaptive")
    P = Nystrom(cfg)
    assert P.U is None and P.S is None
    P._update(_as_linop(spd) if as_linop else spd, DEVICE)
    assert P.U.shape == (50, 20) and P.S.shape == (20,)
    assert P.U.dtype == precision and P.S.dtype == precision and P.U.device == DEVICE
    assert torch.all(P.S >= 0)
    assert torch.all(P.S[:-1] >= P.S[1:])  # descending, so S[-1] is the smallest (adaptive damping uses it)
    assert torch.allclose(P.U.T @ P.U, torch.eye(20, dtype=precision, device=DEVICE), **tol)
    x, X = torch.randn(50, dtype=precision, device=DEVICE), torch.randn(50, 5, dtype=precision, device=DEVICE)
    assert torch.allclose(P @ x, P.U @ (P.S * (P.U.T @ x)) + cfg.rho * x, **tol)
    assert torch.allclose(P @ X, P.U @ (P.S[:, None] * (P.U.T @ X)) + cfg.rho * X, **tol)
    for v in (x, X):
        inv = P._inv @ v
        assert inv.shape == v.shape and inv.dtype == precision
        assert torch.allclose(P @ inv, v, **tol)
    explicit = P.U @ torch.diag(P.S) @ P.U.T + cfg.rho * torch.eye(50, dtype=precision, device=DEVICE)
    assert torch.allclose(explicit @ x, P @ x, **tol)
    # fixed damping ignores the baseline, adaptive adds the smallest Nystrom eigenvalue
    P._update_damping(2.0)
    assert P.config.rho == 1e0
    Pa = Nystrom(NystromConfig(rank=20, sketch=sketch, rho=1e0, damping_mode="adaptive"))
    Pa._update(spd, DEVICE)
    Pa._update_damping(2.0)
    assert torch.isclose(torch.as_tensor(Pa.config.rho, dtype=precision, device=DEVICE), 2.0 + Pa.S[-1])
    assert torch.allclose(Pa @ (Pa._inv @ x), x, **tol)


@pytest.mark.parametrize("as_linop", [False, True], ids=["tensor", "linop"])
def test_newton(spd, precision, as_linop):
    tol = TOL[precision]
    P = Newton(NewtonConfig(rho=1e-1))
    assert P.L is None
    before = spd.clone()
    P._update(_as_linop(spd) if as_linop else spd, DEVICE)
    assert torch.equal(spd, before)  # the caller's matrix is left alone (the reference adds rho in place)
    assert P.L.shape == (50, 50) and P.L.dtype == precision
    M = spd + 1e-1 * torch.eye(50, dtype=precision, device=DEVICE)
    assert torch.allclose(P.L @ P.L.T, M, rtol=tol["rtol"], atol=tol["atol"] * float(M.abs().max()))
    x, X = torch.randn(50, dtype=precision, device=DEVICE), torch.randn(50, 5, dtype=precision, device=DEVICE)
    scale = float(M.abs().max())
    assert torch.allclose(P @ x, M @ x, rtol=tol["rtol"], atol=tol["atol"] * scale)
    for v in (x, X):
        inv = P._inv @ v
        assert inv.shape == v.shape and inv.dtype == precision
        assert torch.allclose(P @ inv, v, rtol=10 * tol["rtol"], atol=100 * tol["atol"])
    assert (x @ (P @ x)) > 0  # SPD


def test_identity(precision):
    P = Identity(IdentityConfig())
    P._update(None, DEVICE)
    x, X = torch.randn(7, dtype=precision), torch.randn(7, 3, dtype=precision)
    for v in (x, X):
        assert torch.equal(P @ v, v) and torch.equal(P._inv @ v, v)


def test_base_class_contract():
    class Doubling(Preconditioner):
        def _update(self, A, device):
            self.seen = device

        def _matmul(self, x):
            return 2 * x

        def _solve(self, x2d):
            return x2d / 2

    P = Doubling(IdentityConfig())
    P._update(None, DEVICE)
    assert P.seen == DEVICE
    x = torch.ones(4)
    with pytest.raises(TypeError):
        P @ [1.0, 2.0]
    with pytest.raises(ValueError):
        P @ torch.ones(2, 2, 2)
    assert torch.equal(P._inv @ x, x / 2) and torch.equal(P._inv @ x[:, None], x[:, None] / 2)
    assert torch.equal(P._inverse_matmul_compose(lambda v: v + 1)(x), (x + 1) / 2)
    assert P._update_damping(1.0) is None
    assert P._inv.preconditioner is P


def test_config_validation():
    with pytest.raises(TypeError):
        NystromConfig(rank=2.5, rho=1.0)
    with pytest.raises(ValueError):
        NystromConfig(rank=0, rho=1.0)
    with pytest.raises(ValueError):
        NystromConfig(rank=5, rho=-1.0)
    with pytest.raises(ValueError):
        NystromConfig(rank=5, rho=1.0, damping_mode="sometimes")
    with pytest.raises(TypeError):
        NewtonConfig(rho=1)
    assert NystromConfig(rank=5, rho=1.0).sketch == "ortho"
    assert NewtonConfig(rho=0.5).to_dict() == {"rho": 0.5}


@pytest.mark.parametrize("name", ["gauss", "ortho"])
def test_sketch_shapes_and_application(name, precision):
    S_right = get_sketch(name, "right", 8, 30, precision, DEVICE)
    S_left = get_sketch(name, "left", 8, 30, precision, DEVICE)
    assert S_right.Omega_mat.shape == (30, 8) and S_left.Omega_mat.shape == (8, 30)
    assert S_right.Omega_mat.is_contiguous() and S_right.Omega_mat.dtype == precision
    M = torch.randn(30, 30, dtype=precision, device=DEVICE)
    assert torch.allclose(S_right._apply_right(M), M @ S_right.Omega_mat)
    assert torch.allclose(S_right._apply_left_trans(M), S_right.Omega_mat.T @ M)
    assert torch.allclose(S_left._apply_left(M), S_left.Omega_mat @ M)
    assert torch.allclose(S_left._apply_right_trans(M), M @ S_left.Omega_mat.T)
    assert torch.allclose(S_right._apply_right(_as_linop(M + M.T)), (M + M.T) @ S_right.Omega_mat, **TOL[precision])
    if name == "ortho":
        assert torch.allclose(S_right.Omega_mat.T @ S_right.Omega_mat, torch.eye(8, dtype=precision, device=DEVICE), **TOL[precision])
        assert torch.allclose(S_left.Omega_mat @ S_left.Omega_mat.T, torch.eye(8, dtype=precision, device=DEVICE), **TOL[precision])
    with pytest.raises(ValueError):
        get_sketch("fourier", "right", 8, 30, precision, DEVICE)
    with pytest.raises(ValueError):
        get_sketch(name, "middle", 8, 30, precision, DEVICE)
    with pytest.raises(NotImplementedError):
        get_sketch("sparse", "right", 8, 30, precision, DEVICE)


def test_reference_import_paths():
    """Module paths a reference user imports from (rlaopt.<pkg>.<module>) resolve to the same objects."""
    import rlaopt_b200.models.linsys as m_linsys
    import rlaopt_b200.preconditioners.configs as p_cfg
    import rlaopt_b200.sketches.gauss as s_gauss
    import rlaopt_b200.solvers.configs as s_cfg
    import rlaopt_b200.solvers.pcg as s_pcg
    import rlaopt_b200.solvers.sap as s_sap
    import rlaopt_b200.spectral_estimators.spectral_norm as sn
    from rlaopt_b200 import models, preconditioners, sketches, solvers, spectral_estimators

    assert m_linsys.LinSys is models.LinSys
    assert p_cfg.NystromConfig is preconditioners.NystromConfig
    assert s_cfg.PCGConfig is solvers.PCGConfig and s_cfg.SAPConfig is solvers.SAPConfig
    assert s_pcg.PCG.__name__ == "PCG" and s_sap.SAP.__name__ == "SAP"
    assert s_gauss.Gauss is sketches.Gauss
    assert sn.randomized_powering is spectral_estimators.randomized_powering
