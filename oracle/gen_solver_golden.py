"""Golden iterates of the reference's OWN solvers (test infrastructure, not product code).

Runs ``rlaopt.models.LinSys.solve`` from the reference with its unmodified ``PCG``, ``SAP``,
``Nystrom``, sketches and ``randomized_powering`` on the CPU, on top of dense-torch kernel
operators built from ``oracle/kernel_oracle.py`` (the reference's own kernel operators need PyKeOps,
which is not installable here -- SURVEY section 8c), and stores the iterates in
``tests/golden/solvers_ref_<dtype>.pt``.

The reference is imported from a scratch build outside the repo::

    cp -r /root/reference /tmp/rlaopt_ref && cd /tmp/rlaopt_ref && \
        RLAOPT_CPU_ONLY=1 RLAOPT_USE_OPENMP=0 python setup.py build_ext --inplace
    python oracle/gen_solver_golden.py float32 && python oracle/gen_solver_golden.py float64

Only this script reads the reference; the tests read the committed fixtures.
"""
import os
import sys

import torch

DTYPE = getattr(torch, sys.argv[1] if len(sys.argv) > 1 else "float32")
torch.set_default_dtype(DTYPE)  # the reference freezes LinOp's default dtype at import (linops/simple.py:12)

REF = os.environ.get("RLAOPT_REF_BUILD", "/tmp/rlaopt_ref")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import rlaopt  # noqa: E402  (the reference)
from rlaopt.linops import LinOp, SymmetricLinOp  # noqa: E402
from rlaopt.models import LinSys  # noqa: E402
from rlaopt.preconditioners import IdentityConfig, NystromConfig  # noqa: E402
from rlaopt.solvers import PCGConfig, SAPAccelConfig, SAPConfig  # noqa: E402
from rlaopt.solvers import sap as ref_sap  # noqa: E402

from oracle import kernel_oracle as ko  # noqa: E402

assert os.path.realpath(rlaopt.__file__).startswith(os.path.realpath(REF)), rlaopt.__file__
CPU = torch.device("cpu")


def dense_system(X, B, reg, kernel, ls):
    """Reference LinSys over a dense kernel matrix with row / block oracles (kernels/base.py:124-128 semantics)."""
    K = ko.kernel_matrix(X, X, kernel, ls, dtype=DTYPE)
    n = X.shape[0]
    A = SymmetricLinOp(CPU, torch.Size((n, n)), lambda v: K @ v, lambda V: K @ V, dtype=DTYPE)

    def row_oracle(blk):
        Kb = K[blk]
        return LinOp(CPU, torch.Size((len(blk), n)), lambda v: Kb @ v, lambda V: Kb @ V, dtype=DTYPE)

    def blk_oracle(blk):
        Kbb = K[blk][:, blk]
        return LinOp(CPU, torch.Size((len(blk), len(blk))), lambda v: Kbb @ v, lambda V: Kbb @ V, dtype=DTYPE)

    return LinSys(A, B, reg=reg, A_row_oracle=row_oracle, A_blk_oracle=blk_oracle)


def snapshot(W, model):
    return W.clone()


def run_case(name, n, d, k, kernel, ls, reg, solver_config, seed, callback_freq, keep):
    g = torch.Generator().manual_seed(seed)
    X = (torch.randn(n, d, generator=g, dtype=torch.float64) / d**0.5).to(DTYPE)
    B = torch.randn(n, k, generator=g, dtype=torch.float64).to(DTYPE)
    system = dense_system(X, B, reg, kernel, ls)
    trace = {"blocks": [], "steps": []}
    if isinstance(solver_config, SAPConfig):  # record the random blocks and step sizes the reference used
        orig_blk, orig_step = ref_sap.SAP._get_blk, ref_sap.SAP._get_stepsize

        def rec_blk(self):
            blk = orig_blk(self)
            trace["blocks"].append(blk.clone())
            return blk

        def rec_step(self, blk, P):
            s = orig_step(self, blk, P)
            trace["steps"].append(float(s))
            return s

        ref_sap.SAP._get_blk, ref_sap.SAP._get_stepsize = rec_blk, rec_step
    torch.manual_seed(seed + 1)  # the solve draws Omega / power-iteration starts / blocks from the global CPU stream
    try:
        W, log = system.solve(solver_config, torch.zeros(n, k, dtype=DTYPE), callback_fn=snapshot,
                              callback_freq=callback_freq)
    finally:
        if isinstance(solver_config, SAPConfig):
            ref_sap.SAP._get_blk, ref_sap.SAP._get_stepsize = orig_blk, orig_step
    iters = sorted(log)
    case = {
        "name": name, "n": n, "d": d, "k": k, "kernel": kernel, "lengthscale": ls, "reg": reg, "seed": seed,
        "callback_freq": callback_freq, "X": X, "B": B, "W_final": W.clone(), "logged_iters": iters,
        "rel_res": torch.stack([log[i]["metrics"]["internal_metrics"]["rel_res"] for i in iters]),
        "W_at": {i: log[i]["metrics"]["callback"] for i in iters if i in keep},
        "blocks": torch.stack(trace["blocks"]) if trace["blocks"] else None,
        "steps": torch.tensor(trace["steps"], dtype=torch.float64) if trace["steps"] else None,
    }
    print(f"{name}: {iters[-1]} iterations, final rel_res {case['rel_res'][-1].tolist()}")
    return case


def main():
    f32 = DTYPE == torch.float32
    rtol = 1e-4 if f32 else 1e-9
    cases = []
    cases.append(run_case(
        "pcg_nystrom_gauss_rbf", 1500, 8, 3, "rbf", 1.0, 0.5,
        PCGConfig(device=CPU, max_iters=60, rtol=rtol, precond_config=NystromConfig(rank=60, rho=0.5, sketch="gauss")),
        seed=0, callback_freq=1, keep={1, 2, 3, 5, 8}))
    cases.append(run_case(
        "pcg_identity_matern52", 1200, 6, 2, "matern52", 1.5, 1.0,
        PCGConfig(device=CPU, max_iters=80, rtol=rtol, precond_config=IdentityConfig()),
        seed=3, callback_freq=1, keep={1, 2, 5, 10}))
    cases.append(run_case(
        "pcg_nystrom_ortho_rbf_k1", 1500, 8, 1, "rbf", 1.0, 0.5,
        PCGConfig(device=CPU, max_iters=60, rtol=rtol, precond_config=NystromConfig(rank=60, rho=0.5)),
        seed=5, callback_freq=1, keep={1, 2, 5}))
    cases.append(run_case(
        "askotch_nystrom_gauss_rbf", 1500, 8, 2, "rbf", 1.0, 0.1,
        SAPConfig(device=CPU, max_iters=60, rtol=rtol, blk_sz=150, precond_config=NystromConfig(rank=30, rho=0.1, sketch="gauss"),
                  accel=True, accel_config=SAPAccelConfig(mu=0.1, nu=10.0), power_iters=10),
        seed=7, callback_freq=10, keep={10, 30, 60}))
    cases.append(run_case(
        "sap_identity_laplace", 900, 5, 1, "laplace", 2.0, 0.2,
        SAPConfig(device=CPU, max_iters=40, rtol=rtol, blk_sz=100, precond_config=IdentityConfig(), accel=False),
        seed=9, callback_freq=10, keep={10, 40}))
    out = os.path.join(ROOT, "tests", "golden", f"solvers_ref_{str(DTYPE).split('.')[-1]}.pt")
    torch.save({"dtype": str(DTYPE), "torch": torch.__version__, "cases": cases}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
