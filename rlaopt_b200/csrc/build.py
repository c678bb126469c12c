"""In-tree build of the C-ABI CUDA library for sm_100a.

    python -m rlaopt_b200.csrc.build [--force]

nvcc cross-compiles without a GPU; the resulting ``librlaopt_b200.so`` is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["kmm_api.cu", "kmm_pack.cu", "kmm_simt.cu", "kmm_tc.cu"]
HEADERS = ["kmm_common.cuh", "kmm_launch.h", "kmm_tmem_ldst.cuh", os.path.join("..", "..", "include", "rlaopt_b200.h")]
OUT = os.path.join(HERE, "librlaopt_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-lcuda",
]


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        path = os.path.join(HERE, name)
        if os.path.exists(path):
            with open(path, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
