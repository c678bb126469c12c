"""Generate the golden fixtures that pin ``oracle/kernel_oracle.py``.

Runs ONLY in the build container (it reads ``/root/reference``); its outputs are
committed under ``tests/golden/`` and travel to the GPU box, where
``/root/reference`` does not exist.

The reference's kernel path cannot be imported (PyKeOps missing), but the
closed-form double loop its own tests use as the pin *can*:
``/root/reference/tests/kernels/utils.py:4-60`` (``compute_kernel_matrix`` and
``rbf_kernel`` / ``laplace_kernel`` / ``matern12|32|52_kernel``).  This script
loads that file by path, evaluates it on seeded inputs that mirror the shapes of
``tests/kernels/test_standard.py:45-130`` (A1 10x3, A2 5x3, blk=[0,1],
const_scaling=2.0, lengthscale 1.0 or [1,2,3], fp32 and fp64) plus one larger
ragged case, and stores inputs and outputs in ``tests/golden/kernels_ref.pt``.

Usage:  python oracle/gen_golden.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
from types import SimpleNamespace

import torch

REF_UTILS = "/root/reference/tests/kernels/utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "kernels_ref.pt")

KERNEL_FUNCS = {
    "rbf": "rbf_kernel",
    "laplace": "laplace_kernel",
    "matern12": "matern12_kernel",
    "matern32": "matern32_kernel",
    "matern52": "matern52_kernel",
}


def _load_reference_utils():
    spec = importlib.util.spec_from_file_location("_ref_kernel_test_utils", REF_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> int:
    if not os.path.exists(REF_UTILS):
        print(f"{REF_UTILS} not found: golden fixtures can only be generated in the build container")
        return 1
    ref = _load_reference_utils()
    gen = torch.Generator().manual_seed(20261018)
    cases = []
    shapes = [
        # (name, n, m, d, blk, lengthscales, const_scaling)
        ("ref_test_shape", 10, 5, 3, [0, 1], [1.0, [1.0, 2.0, 3.0]], 2.0),
        ("ragged", 37, 23, 5, [3, 0, 22, 7, 7 + 4], [0.7, [0.5, 1.0, 1.5, 2.0, 4.0]], 1.0),
    ]
    for name, n, m, d, blk, lengthscales, c in shapes:
        for dtype in (torch.float32, torch.float64):
            A1 = torch.randn(n, d, generator=gen, dtype=torch.float64).to(dtype)
            A2 = torch.randn(m, d, generator=gen, dtype=torch.float64).to(dtype)
            V = torch.randn(m, 2, generator=gen, dtype=torch.float64).to(dtype)
            W = torch.randn(n, 2, generator=gen, dtype=torch.float64).to(dtype)
            blk_t = torch.tensor(blk, dtype=torch.long)
            for ls in lengthscales:
                ls_val = ls if isinstance(ls, float) else torch.tensor(ls, dtype=dtype)
                cfg = SimpleNamespace(const_scaling=c, lengthscale=ls_val)
                for kname, fname in KERNEL_FUNCS.items():
                    fn = getattr(ref, fname)
                    K = ref.compute_kernel_matrix(A1, A2, cfg, torch.device("cpu"), dtype, fn)
                    K_row = ref.compute_kernel_matrix(A1[blk_t], A2, cfg, torch.device("cpu"), dtype, fn)
                    K_blk = ref.compute_kernel_matrix(A1[blk_t], A2[blk_t], cfg, torch.device("cpu"), dtype, fn)
                    cases.append(
                        dict(
                            case=name,
                            kernel=kname,
                            dtype=str(dtype).replace("torch.", ""),
                            const_scaling=c,
                            lengthscale=ls_val,
                            A1=A1,
                            A2=A2,
                            V=V,
                            W=W,
                            blk=blk_t,
                            K=K,  # includes const_scaling (tests/kernels/utils.py:22)
                            K_row=K_row,
                            K_blk=K_blk,
                            KV=K @ V,
                            KtW=K.T @ W,
                        )
                    )
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save({"source": REF_UTILS, "seed": 20261018, "cases": cases}, OUT)
    print(f"wrote {len(cases)} cases to {os.path.relpath(OUT)} ({os.path.getsize(OUT)} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
