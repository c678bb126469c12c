"""Warm timings of the tensor-core pack path (column means, abs-max, pack) on one GPU."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlaopt_b200 import ops
dev = torch.device("cuda:0")
for n, d, gather in ((1_000_000, 128, 0), (10_000_000, 16, 0), (4_000_000, 32, 0), (2_000_000, 64, 100_000)):
    X = torch.randn(n, d, device=dev) / d**0.5
    idx = torch.randperm(n, device=dev)[:gather] if gather else None
    def run():
        c = ops.column_mean(X, idx)
        return ops.pack_points(X, 1.0, idx, ops.LAYOUT_TC, c)
    for _ in range(2): P = run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); c = ops.column_mean(X, idx); b.record(); P = ops.pack_points(X, 1.0, idx, ops.LAYOUT_TC, c); c2.record()
        torch.cuda.synchronize(); ts.append((a.elapsed_time(b), b.elapsed_time(c2)))
    ts.sort(key=lambda t: t[0] + t[1]); m = ts[2]
    print(f"n={n} d={d} gather={gather}: column_mean {m[0]:.3f} ms, absmax+pack {m[1]:.3f} ms, max_sqnorm {P.max_sqnorm:.4f}", flush=True)
    del X, P
