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
for rep in range(4):
    Y = run(mode)
    bad = Y != Y0
    if not bool(bad.any()):
        print("rep", rep, "clean", flush=True)
        continue
    rows = torch.unique(bad.nonzero(as_tuple=True)[0])
    blocks = torch.unique(rows // 128)
    print("rep", rep, "bad row blocks", blocks.numel(), blocks[:12].tolist(), flush=True)
    for rb in blocks[:4].tolist():
        for pair in range(4):
            cs = slice(pair * 256, min((pair + 1) * 256, k))
            sub = bad[rb * 128:(rb + 1) * 128, cs]
            if not bool(sub.any()):
                continue
            D = (Y[rb * 128:(rb + 1) * 128, cs] - Y0[rb * 128:(rb + 1) * 128, cs])          # 128 x 256
            C = (D @ V[:, cs].T) / D.shape[1]                                                # 128 x m: ~ dK[i, m]
            en = (C * C).view(128, m // 64, 64).sum(dim=(0, 2))                              # energy per sub-tile
            top = en.topk(6)
            med = float(en.median())
            print(f"  block {rb} pair {pair}: |D| rms {float(D.pow(2).mean().sqrt()):.3e}; tile energy / median: "
                  + ", ".join(f"t={int(i)} (u={int(i) % 1024}, split {int(i) // 1024}): {float(v) / med:.1f}" for v, i in zip(top.values, top.indices)), flush=True)
            t0 = int(top.indices[0])
            x = A1[rb * 128:(rb + 1) * 128].double()
            Kt = torch.exp(-0.5 * torch.cdist(x, A2[t0 * 64:(t0 + 1) * 64].double()).pow(2))  # 128 x 64
            rel = (C[:, t0 * 64:(t0 + 1) * 64].double() / Kt)
            print("    top tile dK/K by column (mean over rows):", [round(float(v), 3) for v in rel.mean(0)], flush=True)
            print("    top tile dK/K by row (mean over cols, first 16):", [round(float(v), 3) for v in rel.mean(1)[:16]], flush=True)
            del C
    break
