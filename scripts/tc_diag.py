"""Diagnostic sweep of the tcgen05 path against the fp64 oracle (run on the GPU box)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kernel_oracle as ko  # noqa: E402
from rlaopt_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")


def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def run(name, n, m, d, k, layout, scale=1.0, ls=1.0):
    A1 = rnd((n, d), 1) / d**0.5 * scale
    A2 = rnd((m, d), 2) / d**0.5 * scale
    V = rnd((m, k), 3)
    ref = ko.kernel_matmat_gemm_form(A1, A2, V, name, ls, dtype=torch.float64) if name != "laplace" else ko.kernel_matmat(A1, A2, V, name, ls, dtype=torch.float64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = ops.kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), name, ls, layout=layout)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = ko.rel_fro_error(got, ref)
    print(f"{name:9s} n={n:6d} m={m:6d} d={d:3d} k={k:3d} layout={layout} rel_err={err:.3e} max_abs_ref={ref.abs().max():.3e} t={dt*1e3:.1f}ms", flush=True)
    return got, ref


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    TC = _lib.LAYOUT_TC
    if which == "basic":
        # K readout with V = identity: errors are per-entry
        n = m = 64
        A = rnd((n, 16), 5) / 4
        I = torch.eye(m)
        Kref = ko.kernel_matrix(A, A, "rbf", 1.0, dtype=torch.float64)
        Kg = ops.kernel_matmat(A.to(dev), A.to(dev), I.to(dev), "rbf", 1.0, layout=TC).cpu().double()
        print("K readout 64x64 d=16: max abs err", (Kg - Kref).abs().max().item(), "diag", Kg.diagonal()[:4].tolist(), flush=True)
        if (Kg - Kref).abs().max().item() > 1e-3:
            print("Kg[:4,:4]\n", Kg[:4, :4], "\nKref[:4,:4]\n", Kref[:4, :4], flush=True)
        for args in [
            ("rbf", 128, 64, 16, 16),
            ("rbf", 128, 64, 64, 64),
            ("rbf", 128, 128, 128, 64),
            ("rbf", 100, 300, 3, 1),
            ("rbf", 1000, 2000, 128, 64),
            ("matern32", 1000, 2000, 32, 16),
            ("matern52", 1000, 2000, 32, 16),
            ("matern12", 1000, 2000, 32, 16),
            ("rbf", 300, 5000, 100, 130),
            ("rbf", 300, 5000, 192, 64),
            ("rbf", 70, 40000, 50, 7),
            ("rbf", 8192, 8192, 128, 64),
        ]:
            run(*args, TC)
        run("rbf", 8192, 8192, 128, 64, _lib.LAYOUT_SIMT)
    elif which == "big":
        for args in [("rbf", 32768, 32768, 128, 64), ("matern52", 32768, 32768, 32, 16), ("rbf", 65536, 65536, 128, 64)]:
            run(*args, TC)
            run(*args, TC)
    elif which == "period":
        for per in ("1", "2", "4", "8", "16", "64"):
            os.environ["RLAOPT_B200_TC_PERIOD"] = per
            print("PERIOD", per, flush=True)
            run("rbf", 2048, 16384, 128, 64, TC)
            run("matern52", 2048, 16384, 32, 16, TC)
        os.environ.pop("RLAOPT_B200_TC_PERIOD")
        # all-positive V (coherent sums): bias shows up as a relative offset
        A = rnd((2048, 128), 1) / 128**0.5
        B = rnd((16384, 128), 2) / 128**0.5
        V = rnd((16384, 64), 3).abs()
        ref = ko.kernel_matmat_gemm_form(A, B, V, "rbf", 1.0, dtype=torch.float64)
        for per in ("1", "4", "16"):
            os.environ["RLAOPT_B200_TC_PERIOD"] = per
            got = ops.kernel_matmat(A.to(dev), B.to(dev), V.to(dev), "rbf", 1.0, layout=TC).cpu().double()
            rel = ((got - ref) / ref)
            print(f"positive V period {per}: mean rel {rel.mean().item():.3e} rms {rel.pow(2).mean().sqrt().item():.3e}", flush=True)
    elif which == "perf":
        from rlaopt_b200.kernels import KernelConfig, RBFLinOp, Matern52LinOp
        for per in (("4", "2"), ("3", "1"), ("2", "1")):
            os.environ["RLAOPT_B200_TC_NB"], os.environ["RLAOPT_B200_TC_LA"] = per
            for cls, n, d, k in ((RBFLinOp, 131072, 128, 64), (Matern52LinOp, 262144, 32, 16)):
                X = (rnd((n, d), 1) / d**0.5).to(dev)
                V = rnd((n, k), 2).to(dev)
                op = cls(X, X, KernelConfig(lengthscale=1.0))
                for _ in range(2):
                    Y = op @ V
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    Y = op @ V
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                print(f"NB,LA={per} {cls.__name__} n={n} d={d} k={k}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s", flush=True)
    elif which == "knock":
        from rlaopt_b200.kernels import KernelConfig, RBFLinOp, Matern52LinOp
        for diag in ("0", "1", "2", "4", "6", "8", "9", "3", "15"):
            os.environ["RLAOPT_B200_TC_DIAG"] = diag
            for cls, n, d, k in ((RBFLinOp, 131072, 128, 64), (Matern52LinOp, 262144, 32, 16)):
                X = (rnd((n, d), 1) / d**0.5).to(dev)
                V = rnd((n, k), 2).to(dev)
                op = cls(X, X, KernelConfig(lengthscale=1.0))
                for _ in range(2):
                    Y = op @ V
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    Y = op @ V
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                print(f"diag {diag} {cls.__name__} n={n} d={d} k={k}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s", flush=True)
    elif which == "prof2":
        # per-section cycle counters of CTA 0 (needs the -DKMM_TC_PROFILE build, RLAOPT_B200_LIB=...)
        import ctypes
        from rlaopt_b200.kernels import KernelConfig
        lib = _lib.load()
        from rlaopt_b200 import kernels as K
        cls_name = sys.argv[2] if len(sys.argv) > 2 else "RBFLinOp"
        n, d, k = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (131072, 128, 64)
        X = (rnd((n, d), 1) / d**0.5).to(dev)
        V = rnd((n, k), 2).to(dev)
        op = getattr(K, cls_name)(X, X, KernelConfig(lengthscale=1.0))
        print(cls_name, n, d, k, "NB", os.environ.get("RLAOPT_B200_TC_NB"), flush=True)
        for diag in ("0", "2", "1", "3"):
            os.environ["RLAOPT_B200_TC_DIAG"] = diag
            for _ in range(2):
                Y = op @ V
            torch.cuda.synchronize()
            buf = (ctypes.c_longlong * 64)()
            lib.kmm_tc_prof_read(buf)
            tiles = n // 64
            names_m = ["w_a_full", "w_p_free", "issue1+commit", "w_p_full", "w_v_full", "w_o_free", "issue2+commit", "other"]
            names_e = ["w_v_full", "w_s_full", "ld S", "compute+st", "wait st+arrive", "drain", "-", "other"]
            print(f"diag {diag}: MMA1 warp clk/tile:", {nm: round(buf[i] / tiles, 1) for i, nm in enumerate(names_m)}, "total", round(sum(buf[0:8]) / tiles, 1), flush=True)
            print(f"   MMA2 warp clk/tile:", {nm: round(buf[24 + i] / tiles, 1) for i, nm in enumerate(names_m)}, "total", round(sum(buf[24:32]) / tiles, 1), flush=True)
            for base, nm in ((8, "epi warp0"), (16, "epi warp4")):
                print(f"   {nm} clk per own tile:", {x: round(buf[base + i] / (tiles / 2), 1) for i, x in enumerate(names_e)}, "total", round(sum(buf[base:base + 8]) / (tiles / 2), 1), flush=True)
    elif which == "sustain":
        # power-limited regime: ~4 s of back-to-back launches per knock-out, throughput over the last half
        import subprocess
        from rlaopt_b200.kernels import KernelConfig, RBFLinOp
        n, d, k = 262144, 128, 64
        X = (rnd((n, d), 1) / d**0.5).to(dev)
        V = rnd((n, k), 2).to(dev)
        op = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
        for diag in (sys.argv[2:] or ["0", "8", "2", "4", "1"]):
            os.environ["RLAOPT_B200_TC_DIAG"] = diag
            Y = op @ V
            torch.cuda.synchronize()
            reps = 30
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps + 1)]
            ev[0].record()
            for i in range(2 * reps):
                Y = op @ V
                ev[i + 1].record()
            smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
            torch.cuda.synchronize()
            ms = ev[reps].elapsed_time(ev[2 * reps]) / reps
            print(f"sustain diag {diag}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s  [{smi}]", flush=True)
    elif which == "kinds":
        # throughput of every kernel family at the C3 shape (d=32, k=16) and the C2 shape, one GPU
        from rlaopt_b200 import kernels as K
        from rlaopt_b200.kernels import KernelConfig
        for name, n, d, k in (("LaplaceLinOp", 131072, 32, 16), ("Matern12LinOp", 131072, 32, 16), ("Matern32LinOp", 262144, 32, 16),
                              ("Matern52LinOp", 262144, 32, 16), ("RBFLinOp", 262144, 32, 16), ("LaplaceLinOp", 65536, 128, 64),
                              ("RBFLinOp", 131072, 64, 128), ("RBFLinOp", 131072, 64, 1000), ("RBFLinOp", 131072, 16, 1), ("RBFLinOp", 131072, 8, 10), ("RBFLinOp", 131072, 128, 16), ("RBFLinOp", 131072, 100, 1), ("RBFLinOp", 65536, 256, 16), ("RBFLinOp", 65536, 784, 1), ("Matern52LinOp", 32768, 1024, 64), ("RBFLinOp", 32768, 784, 200)):
            X = (rnd((n, d), 1) / d**0.5).to(dev)
            V = rnd((n, k), 2).to(dev)
            op = getattr(K, name)(X, X, KernelConfig(lengthscale=1.0))
            for _ in range(2):
                Y = op @ V
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                Y = op @ V
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"{name} n={n} d={d} k={k}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s", flush=True)
    elif which == "l2chunk":
        # column-chunked launch order: DRAM traffic vs L2-sized chunks, power-limited regime
        import subprocess
        from rlaopt_b200.kernels import KernelConfig, RBFLinOp
        n, d, k = 524288, 128, 64
        X = (rnd((n, d), 1) / d**0.5).to(dev)
        V = rnd((n, k), 2).to(dev)
        for cap in (sys.argv[2:] or ["0", "2048", "1024", "512", "256"]):
            os.environ["RLAOPT_B200_TC_SPLIT_TILES"] = cap
            op = RBFLinOp(X, X, KernelConfig(lengthscale=1.0))
            Y = op @ V
            torch.cuda.synchronize()
            reps = 8
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps + 1)]
            ev[0].record()
            for i in range(2 * reps):
                Y = op @ V
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = ev[reps].elapsed_time(ev[2 * reps]) / reps
            print(f"split cap {cap}: {ms:.2f} ms  {n*n/ms/1e6:.1f} Gentries/s checksum {float(Y.double().abs().sum()):.6e}", flush=True)
    elif which == "range":
        # tiny kernel values (far-apart clusters): the per-row power-of-two scale keeps relative accuracy
        for shift in (0.0, 0.5, 1.0, 1.5):
            n, m, d, k = 512, 4096, 64, 32
            A1 = rnd((n, d), 1) / d**0.5
            A2 = rnd((m, d), 2) / d**0.5 + shift
            V = rnd((m, k), 3)
            ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
            for lay in (TC, _lib.LAYOUT_SIMT):
                got = ops.kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=lay)
                print(f"shift {shift}: layout {lay} rel_err {ko.rel_fro_error(got, ref):.3e} |ref|max {ref.abs().max():.3e}", flush=True)
        # V with a huge dynamic range across row tiles
        n, m, d, k = 512, 4096, 64, 32
        A1 = rnd((n, d), 1) / d**0.5
        A2 = rnd((m, d), 2) / d**0.5
        V = rnd((m, k), 3) * torch.logspace(-12, 12, m).unsqueeze(1)
        ref = ko.kernel_matmat_gemm_form(A1, A2, V, "rbf", 1.0, dtype=torch.float64)
        for lay in (TC, _lib.LAYOUT_SIMT):
            got = ops.kernel_matmat(A1.to(dev), A2.to(dev), V.to(dev), "rbf", 1.0, layout=lay)
            print(f"wide V: layout {lay} rel_err {ko.rel_fro_error(got, ref):.3e}", flush=True)
    elif which == "prof":
        from rlaopt_b200.kernels import KernelConfig, RBFLinOp
        n, m, d, k = 148 * 128, 65536, 128, 64
        X1 = (rnd((n, d), 1) / d**0.5).to(dev)
        X2 = (rnd((m, d), 2) / d**0.5).to(dev)
        V = rnd((m, k), 3).to(dev)
        op = RBFLinOp(X1, X2, KernelConfig(lengthscale=1.0))
        for _ in range(3):
            Y = op @ V
        torch.cuda.synchronize()
        print("done", float(Y.abs().sum()))
