"""Which sub-tile is wrong when the sliced two-chunk kernel mis-computes a row block?  D = Y - Y0 of a bad row is a
combination of the rows of V of the bad column tile: correlate it with V (random, so V V^T ~ k I)."""
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
    f = mode.split(":")
    os.environ["RLAOPT_B200_TC_DUAL"] = f[0]
    os.environ["RLAOPT_B200_TC_DUAL_OVERLAP"] = f[1] if len(f) > 1 else "1"
    Y = op @ V
    torch.cuda.synchronize()
    return Y
Y0 = run("0")
mode = sys.argv[1] if len(sys.argv) > 1 else "3:0"
found = 0
for rep in range(6):
    Y = run(mode)
    bad = Y != Y0
    if not bool(bad.any()):
        print("rep", rep, "clean", flush=True)
        continue
    rows = torch.unique(bad.nonzero(as_tuple=True)[0])
    blocks = torch.unique(rows // 128)
    print("rep", rep, "bad row blocks", blocks.numel(), blocks[:40].tolist(), flush=True)
    for rb in blocks[:16].tolist():
        for pair in range(4):
            cs = slice(pair * 256, min((pair + 1) * 256, k))
            sub = bad[rb * 128:(rb + 1) * 128, cs]
            if not bool(sub.any()):
                continue
            D = (Y[rb * 128:(rb + 1) * 128, cs] - Y0[rb * 128:(rb + 1) * 128, cs])
            C = (D @ V[:, cs].T) / D.shape[1]
            en = (C * C).view(128, m // 64, 64).sum(dim=(0, 2))
            top = en.topk(3)
            med = float(en.median())
            t0 = int(top.indices[0])
            u, sp = t0 % 1024, t0 // 1024
            T = min(1024, m // 64 - sp * 1024)
            # which rows / columns of the bad tile carry the error
            Ct = C[:, t0 * 64:(t0 + 1) * 64]
            rowe = (Ct * Ct).sum(1); cole = (Ct * Ct).sum(0)
            print(f"  block {rb} pair {pair}: tile {t0} = split {sp} u {u} of T {T} (u%2 {u % 2} u%3 {u % 3} u%4 {u % 4} u%8 {u % 8}, T-u {T - u}) "
                  f"energy/median {float(top.values[0]) / med:.1f} next {float(top.values[1]) / med:.1f}; "
                  f"row quarters {[round(float(rowe[q * 32:(q + 1) * 32].sum() / rowe.sum()), 2) for q in range(4)]} "
                  f"col quarters {[round(float(cole[q * 16:(q + 1) * 16].sum() / cole.sum()), 2) for q in range(4)]} "
                  f"|D| chunkA {float(D[:, :128].abs().mean()):.2e} chunkB {float(D[:, 128:].abs().mean()):.2e}", flush=True)
            del C
            found += 1
    if found >= 24:
        break
