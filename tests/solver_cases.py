"""Shared helpers for the solver parity tests: rebuild the golden cases with this package's stack."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases(dtype_name):
    blob = torch.load(os.path.join(GOLDEN_DIR, f"solvers_ref_{dtype_name}.pt"), weights_only=False)
    return {c["name"]: c for c in blob["cases"]}


def solver_config_for(name, device, rtol):
    """The configs of oracle/gen_solver_golden.py, expressed with rlaopt_b200's classes."""
    from rlaopt_b200.preconditioners import IdentityConfig, NystromConfig
    from rlaopt_b200.solvers import PCGConfig, SAPAccelConfig, SAPConfig

    if name == "pcg_nystrom_gauss_rbf":
        return PCGConfig(device=device, max_iters=60, rtol=rtol,
                         precond_config=NystromConfig(rank=60, rho=0.5, sketch="gauss"))
    if name == "pcg_identity_matern52":
        return PCGConfig(device=device, max_iters=80, rtol=rtol, precond_config=IdentityConfig())
    if name == "pcg_nystrom_ortho_rbf_k1":
        return PCGConfig(device=device, max_iters=60, rtol=rtol, precond_config=NystromConfig(rank=60, rho=0.5))
    if name == "askotch_nystrom_gauss_rbf":
        return SAPConfig(device=device, max_iters=60, rtol=rtol, blk_sz=150,
                         precond_config=NystromConfig(rank=30, rho=0.1, sketch="gauss"), accel=True,
                         accel_config=SAPAccelConfig(mu=0.1, nu=10.0), power_iters=10)
    if name == "sap_identity_laplace":
        return SAPConfig(device=device, max_iters=40, rtol=rtol, blk_sz=100, precond_config=IdentityConfig(),
                         accel=False)
    raise KeyError(name)


def dense_linsys(case, dtype):
    """LinSys over a dense CPU kernel matrix from the oracle (host-logic tests; no GPU, no CUDA kernels)."""
    from oracle import kernel_oracle as ko
    from rlaopt_b200.linops import LinOp, SymmetricLinOp
    from rlaopt_b200.models import LinSys

    cpu = torch.device("cpu")
    X, B, n = case["X"], case["B"], case["n"]
    K = ko.kernel_matrix(X, X, case["kernel"], case["lengthscale"], dtype=dtype)
    A = SymmetricLinOp(cpu, torch.Size((n, n)), lambda v: K @ v, lambda V: K @ V, dtype=dtype)

    def row_oracle(blk):
        Kb = K[blk]
        return LinOp(cpu, torch.Size((len(blk), n)), lambda v: Kb @ v, lambda V: Kb @ V, dtype=dtype)

    def blk_oracle(blk):
        Kbb = K[blk][:, blk]
        return LinOp(cpu, torch.Size((len(blk), len(blk))), lambda v: Kbb @ v, lambda V: Kbb @ V, dtype=dtype)

    return LinSys(A, B, reg=case["reg"], A_row_oracle=row_oracle, A_blk_oracle=blk_oracle)


def kernel_linsys(case, device):
    """LinSys over this package's fused kernel operators on ``device`` (the product path)."""
    from rlaopt_b200 import kernels
    from rlaopt_b200.kernels import KernelConfig
    from rlaopt_b200.models import LinSys

    cls = {"rbf": kernels.RBFLinOp, "laplace": kernels.LaplaceLinOp, "matern12": kernels.Matern12LinOp,
           "matern32": kernels.Matern32LinOp, "matern52": kernels.Matern52LinOp}[case["kernel"]]
    X, B = case["X"].to(device), case["B"].to(device)
    A = cls(X, X, KernelConfig(lengthscale=float(case["lengthscale"])))
    return LinSys(A, B, reg=case["reg"], A_row_oracle=A.row_oracle, A_blk_oracle=A.blk_oracle)


class Recorder:
    """callback_fn that keeps W at selected iterations, plus hooks recording SAP's blocks and step sizes."""

    def __init__(self):
        self.W = []
        self.blocks, self.steps = [], []

    def __call__(self, W, model):
        self.W.append(W.detach().clone())
        return None

    def hook(self):
        from rlaopt_b200.solvers import SAP

        rec = self
        orig_blk, orig_step = SAP._get_blk, SAP._get_stepsize

        def blk(self_):
            b = orig_blk(self_)
            rec.blocks.append(b.clone())
            return b

        def step(self_, *a, **k):
            s = orig_step(self_, *a, **k)
            rec.steps.append(float(s))
            return s

        SAP._get_blk, SAP._get_stepsize = blk, step
        return lambda: (setattr(SAP, "_get_blk", orig_blk), setattr(SAP, "_get_stepsize", orig_step))
