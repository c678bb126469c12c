"""In-tree build of the C-ABI CUDA library for sm_100a.

    python -m rlaopt_b200.csrc.build [--force] [-v]

nvcc cross-compiles without a GPU; the resulting ``librlaopt_b200.so`` is
git-ignored but travels to the GPU box with the repo snapshot.  Every ``.cu`` is
compiled to its own object (in parallel, re-done only when the file, a header or
the flags change) and the objects are linked into the shared library.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["kmm_api.cu", "kmm_pack.cu", "kmm_simt.cu", "kmm_tc.cu", "kmm_fuse.cu"]
HEADERS = ["kmm_common.cuh", "kmm_launch.h", "kmm_tmem_ldst.cuh", os.path.join("..", "..", "include", "rlaopt_b200.h")]
OUT = os.path.join(HERE, "librlaopt_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-lcuda"]


def _nvcc() -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return nvcc if os.path.exists(nvcc) else "nvcc"


def _sources() -> list[str]:
    return [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def _file_digest(names: list[str], extra: str = "") -> str:
    h = hashlib.sha256()
    for name in names:
        path = os.path.join(HERE, name)
        if os.path.exists(path):
            with open(path, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(extra.encode())
    return h.hexdigest()


def _digest() -> str:
    return _file_digest(_sources() + HEADERS, " ".join(NVCC_FLAGS + LINK_FLAGS))


def _compile_one(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".stamp"
    digest = _file_digest([src] + HEADERS, " ".join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return obj
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    with open(stamp, "w") as f:
        f.write(digest)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for name in os.listdir(OBJ_DIR):
            if name.endswith(".stamp"):
                os.remove(os.path.join(OBJ_DIR, name))
    with ThreadPoolExecutor(max_workers=min(8, len(_sources()))) as pool:
        objs = list(pool.map(lambda s: _compile_one(s, verbose), _sources()))
    cmd = [_nvcc()] + ["-gencode", "arch=compute_100a,code=sm_100a"] + LINK_FLAGS + ["-Xcompiler", "-fPIC", "-o", OUT] + objs
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc link failed ({proc.returncode}): {' '.join(cmd)}")
    with open(STAMP, "w") as f:
        f.write(digest)
    return OUT


# ---- torch.library front (torch.ops.rlaopt.kernel_matmat): host C++ only, links the C-ABI library ----------------
TORCH_OP_SRC = "torch_op.cpp"
TORCH_OP_OUT = os.path.join(HERE, "librlaopt_b200_torch.so")
TORCH_OP_STAMP = os.path.join(HERE, ".build_stamp_torch")


def build_torch_op(force: bool = False) -> str:
    """g++ torch_op.cpp -> librlaopt_b200_torch.so (registers TORCH_LIBRARY_FRAGMENT(rlaopt, ...); needs the C-ABI
    library built first).  Rebuilt when the source, the header or the torch version changes."""
    import torch
    from torch.utils import cpp_extension

    build(force=False)
    digest = _file_digest([TORCH_OP_SRC, HEADERS[-1]], torch.__version__)
    if not force and os.path.exists(TORCH_OP_OUT) and os.path.exists(TORCH_OP_STAMP):
        with open(TORCH_OP_STAMP) as f:
            if f.read().strip() == digest:
                return TORCH_OP_OUT
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
    tlib = cpp_extension.library_paths()[0]
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in cpp_extension.include_paths()] + ["-I/usr/local/cuda/include"]
    cmd += [TORCH_OP_SRC, "-o", TORCH_OP_OUT, f"-L{HERE}", "-lrlaopt_b200", f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10",
            "-ltorch_cuda", "-lc10_cuda", "-Wl,-rpath,$ORIGIN"]
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"g++ failed ({proc.returncode}): {' '.join(cmd)}")
    with open(TORCH_OP_STAMP, "w") as f:
        f.write(digest)
    return TORCH_OP_OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
    print(build_torch_op(force="--force" in sys.argv))
