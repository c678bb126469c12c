"""Register-contraction mode of the tensor-core kernel (k <= 4): accuracy against an fp64 dense product and
throughput with the mode on / off (RLAOPT_B200_TC_KV).

    python scripts/_kv_check.py [acc|perf]
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200._lib import LAYOUT_TC
from rlaopt_b200.ops import kernel_matmat

dev = torch.device("cuda:0")


def dense64(A1, A2, V, name):
    D2 = torch.cdist(A1.double(), A2.double()).square()
    r = D2.sqrt()
    if name == "rbf":
        K = torch.exp(-0.5 * D2)
    elif name == "matern12":
        K = torch.exp(-r)
    elif name == "matern32":
        K = (1 + 3**0.5 * r) * torch.exp(-(3**0.5) * r)
    else:
        K = (1 + 5**0.5 * r + 5.0 / 3.0 * D2) * torch.exp(-(5**0.5) * r)
    return K @ V.double()


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


def acc():
    worst = 0.0
    for name in ("rbf", "matern12", "matern32", "matern52"):
        for (n, m, d, k) in [(300, 1000, 8, 1), (2048, 4099, 16, 1), (1000, 777, 32, 2), (515, 2100, 100, 3),
                             (130, 64 * 40 + 1, 150, 4), (4096, 70000, 16, 1)]:
            g = torch.Generator().manual_seed(n + m + d + k)
            A1 = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
            A2 = (torch.randn(m, d, generator=g) / d**0.5).to(dev)
            V = torch.randn(m, k, generator=g).to(dev)
            ref = dense64(A1, A2, V, name)
            os.environ["RLAOPT_B200_TC_KV"] = "1"
            got = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC).reshape(n, -1)
            os.environ["RLAOPT_B200_TC_KV"] = "0"
            old = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC).reshape(n, -1)
            os.environ["RLAOPT_B200_TC_KV"] = "1"
            e, e0 = rel(got, ref), rel(old, ref)
            worst = max(worst, e)
            print(f"{name:9s} n={n} m={m} d={d} k={k}: rel err {e:.2e} (MMA2 path {e0:.2e})", flush=True)
    # symmetric operator with duplicates: Matern-1/2 diagonal
    g = torch.Generator().manual_seed(5)
    X = (torch.randn(3000, 16, generator=g) / 4).to(dev)
    V = torch.randn(3000, 1, generator=g).to(dev)
    e = rel(kernel_matmat(X, X, V, "matern12", 1.0, layout=LAYOUT_TC), dense64(X, X, V, "matern12"))
    print(f"matern12 K(X,X) k=1: rel err {e:.2e}")
    worst = max(worst, e)
    # positive V, long sum
    X = (torch.randn(512, 8, generator=g) / 8**0.5).to(dev)
    Z = (torch.randn(400000, 8, generator=g) / 8**0.5).to(dev)
    V = torch.rand(400000, 1, generator=g).to(dev)
    e = rel(kernel_matmat(X, Z, V, "rbf", 1.0, layout=LAYOUT_TC), dense64(X, Z, V, "rbf"))
    print(f"rbf long positive sum m=400000: rel err {e:.2e}")
    worst = max(worst, e)
    print("worst", worst, "OK" if worst <= 2e-6 else "FAIL")


def sweep():
    """ring depth / S buffers of the register-contraction mode"""
    name, n, m, d, k = "rbf", 131072, 1 << 20, 16, 1
    g = torch.Generator().manual_seed(0)
    A1 = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
    A2 = (torch.randn(m, d, generator=g) / d**0.5).to(dev)
    V = torch.randn(m, k, generator=g).to(dev)
    for env in [{"RLAOPT_B200_TC_SA": "4"}, {"RLAOPT_B200_TC_SA": "6"}, {"RLAOPT_B200_TC_SA": "8"}, {"RLAOPT_B200_TC_SA": "10"},
                {"RLAOPT_B200_TC_SA": "8", "RLAOPT_B200_TC_NB": "4"}, {"RLAOPT_B200_TC_SA": "8", "RLAOPT_B200_TC_LA": "1"},
                {"RLAOPT_B200_TC_SA": "8", "RLAOPT_B200_TC_LA": "4"}, {"RLAOPT_B200_TC_SA": "8", "RLAOPT_B200_TC_NWG": "3"}]:
        for key in ("RLAOPT_B200_TC_SA", "RLAOPT_B200_TC_NB", "RLAOPT_B200_TC_LA", "RLAOPT_B200_TC_NWG"):
            os.environ.pop(key, None)
        os.environ.update(env)
        Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{env}: {n * m / sorted(ts)[1] / 1e6:7.1f} Gentries/s", flush=True)


def mid():
    """pointwise-bound family (k = 16, d = 32) and the C2 tile shape, all kernels"""
    for name, n, m, d, k in [("rbf", 131072, 1 << 20, 32, 16), ("matern12", 131072, 1 << 20, 32, 16),
                             ("matern32", 131072, 1 << 20, 32, 16), ("matern52", 131072, 1 << 20, 32, 16),
                             ("rbf", 131072, 1 << 20, 8, 10), ("matern52", 131072, 1 << 20, 64, 32),
                             ("rbf", 131072, 131072, 128, 64), ("matern52", 131072, 131072, 128, 64)]:
        g = torch.Generator().manual_seed(0)
        A1 = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
        A2 = (torch.randn(m, d, generator=g) / d**0.5).to(dev)
        V = torch.randn(m, k, generator=g).to(dev)
        Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{name:9s} n={n} m={m} d={d} k={k}: {n * m / sorted(ts)[1] / 1e6:7.1f} Gentries/s", flush=True)


def wide():
    """wide-d instantiations (193 <= d <= 2048): throughput and agreement with the CUDA-core kernel"""
    from rlaopt_b200._lib import LAYOUT_SIMT
    for name, n, m, d, k in [("rbf", 32768, 262144, 256, 16), ("matern12", 32768, 262144, 256, 16),
                             ("matern32", 32768, 262144, 256, 16), ("matern52", 32768, 262144, 256, 16),
                             ("rbf", 32768, 262144, 784, 1), ("matern52", 16384, 131072, 1024, 64)]:
        g = torch.Generator().manual_seed(0)
        A1 = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
        A2 = (torch.randn(m, d, generator=g) / d**0.5).to(dev)
        V = torch.randn(m, k, generator=g).to(dev)
        Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        R = kernel_matmat(A1[:2048], A2, V, name, 1.0, layout=LAYOUT_SIMT)
        print(f"{name:9s} n={n} m={m} d={d} k={k}: {n * m / sorted(ts)[1] / 1e6:7.1f} Gentries/s, "
              f"rel. diff to the CUDA-core kernel on 2048 rows {rel(Y[:2048], R.double()):.2e}", flush=True)


def perf():
    for name, n, m, d, k in [("rbf", 131072, 1 << 20, 16, 1), ("rbf", 131072, 1 << 20, 8, 1), ("rbf", 131072, 1 << 20, 32, 2),
                             ("rbf", 131072, 1 << 20, 32, 4), ("matern52", 131072, 1 << 20, 16, 1),
                             ("matern12", 131072, 1 << 20, 16, 1), ("rbf", 65536, 1 << 20, 128, 1)]:
        g = torch.Generator().manual_seed(0)
        A1 = (torch.randn(n, d, generator=g) / d**0.5).to(dev)
        A2 = (torch.randn(m, d, generator=g) / d**0.5).to(dev)
        V = torch.randn(m, k, generator=g).to(dev)
        out = []
        variants = [{"RLAOPT_B200_TC_KV": "0"}, {"RLAOPT_B200_TC_KV": "1"}]
        for env in variants:
            for key in ("RLAOPT_B200_TC_KV", "RLAOPT_B200_TC_NWG", "RLAOPT_B200_TC_POLY"):
                os.environ.pop(key, None)
            os.environ.update(env)
            Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                Y = kernel_matmat(A1, A2, V, name, 1.0, layout=LAYOUT_TC)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out.append(n * m / sorted(ts)[1] / 1e6)
        for key in ("RLAOPT_B200_TC_KV", "RLAOPT_B200_TC_NWG", "RLAOPT_B200_TC_POLY"):
            os.environ.pop(key, None)
        print(f"{name:9s} n={n} m={m} d={d} k={k}: MMA2 path {out[0]:7.1f} | register contraction {out[1]:7.1f} Gentries/s "
              f"({out[1] / out[0]:.2f}x) (includes packing X and V per call)", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "acc"
    {"acc": acc, "perf": perf, "sweep": sweep, "mid": mid, "wide": wide}[what]()
