"""Bitwise comparison of the two-chunk kernels against the one-chunk kernel (same arithmetic, same summation order)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import kernels as K
from rlaopt_b200.kernels import KernelConfig

dev = torch.device("cuda:0")
n, m, d, k = 94720, 2_000_000, 64, 1000
g = torch.Generator(device=dev).manual_seed(0)
A2 = torch.randn(m, d, generator=g, device=dev) / d**0.5
A1 = A2[:n]
V = torch.randn(m, k, generator=g, device=dev)
op = K.RBFLinOp(A1, A2, KernelConfig(lengthscale=1.0))
def run(mode):
    # mode = "<DUAL>[:<overlap bits>[:<pair>]]"
    f = mode.split(":")
    os.environ["RLAOPT_B200_TC_DUAL"] = f[0]
    os.environ["RLAOPT_B200_TC_DUAL_OVERLAP"] = f[1] if len(f) > 1 else "1"
    os.environ["RLAOPT_B200_TC_PAIR"] = f[2] if len(f) > 2 else "-1"
    if len(f) > 3:
        os.environ["RLAOPT_B200_TC_SA"] = f[3]  # A-ring depth
    else:
        os.environ.pop("RLAOPT_B200_TC_SA", None)
    Y = op @ V
    torch.cuda.synchronize()
    return Y
Y0 = run("0")
print("mode 0 repeat identical:", bool((run("0") == Y0).all()), flush=True)
for mode in sys.argv[1:] or ["3:0", "3:0", "3:0", "3:0", "3:0", "3:0", "3:1", "3:1", "3:1"]:
    Y = run(mode)
    bad = (Y != Y0)
    nb = int(bad.sum())
    msg = f"mode {mode}: mismatching entries {nb}"
    if nb:
        r, c = bad.nonzero(as_tuple=True)
        rel = ((Y[bad] - Y0[bad]).abs() / Y0[bad].abs().clamp_min(1e-30))
        rows = torch.unique(r); cols = torch.unique(c)
        msg += (f"; rows {rows.numel()} (first {rows[:8].tolist()}, mod 128 {torch.unique(rows % 128)[:16].tolist()}), cols {cols.numel()} "
                f"(first {cols[:8].tolist()}, chunks {torch.unique(cols // 128).tolist()}, col%128//32 {torch.unique(cols % 128 // 32).tolist()}); "
                f"rel diff max {float(rel.max()):.3e} median {float(rel.median()):.3e}")
    print(msg, flush=True)
