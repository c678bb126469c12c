"""Find the first non-finite quantity in a block-PCG solve whose columns converge at different iterations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200.kernels import KernelConfig, RBFLinOp
from rlaopt_b200.models import LinSys
from rlaopt_b200.preconditioners import NystromConfig
from rlaopt_b200.solvers import PCGConfig
from rlaopt_b200.solvers import _pcg

dev = torch.device("cuda:0")
n, rank, reg, mode = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
d, k = 64, 16
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(n, d, generator=g, device=dev) / d**0.5
B = torch.randn(n, k, generator=g, device=dev)
A = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
system = LinSys(A, B, reg=reg)
cfg = PCGConfig(device=dev, max_iters=80, rtol=1e-3, atol=1e-30, precond_config=NystromConfig(rank=rank, rho=reg, sketch="gauss"))
orig_step = _pcg.PCG._step
state = {"it": 0}
def step(self):
    state["it"] += 1
    m = self.system.mask
    orig_step(self)
    bad = [nm for nm, t in (("W", self._W), ("R", self.R), ("Z", self.Z), ("P_", self.P_), ("RZ", self.RZ)) if not bool(torch.isfinite(t).all())]
    print(f"  step {state['it']:3d} active {int(m.sum()):2d} |R| max {float(self.R.norm(dim=0).max()):.3e} RZ diag min {float(self.RZ.diagonal().abs().min()):.3e} cond(RZ_act) "
          f"{float(torch.linalg.cond(self.RZ[m.to(dev)][:, m.to(dev)].double())) if m.any() else 0:.2e} bad {bad}")
_pcg.PCG._step = step
torch.manual_seed(0)
W, log = system.solve(cfg, torch.zeros(n, k, device=dev), callback_freq=1, residual=mode)
print("iterations", max(log), "final", float(log[max(log)]["metrics"]["internal_metrics"]["rel_res"].max()))
