"""Host-side operator layer over the C ABI (``include/rlaopt_b200.h``).

Three levels, all CUDA-only (no CPU fallback — a CPU tensor raises):

* :func:`pack_points` / :func:`matmat_packed` — pack an operand once, apply many
  times (what the LinOp classes use; the pack replaces the LazyTensor the
  reference builds at ``rlaopt/kernels/base.py:88-102``).
* :func:`kernel_matmat` — one call = gather + pack + fused matmat (the whole of
  ``_KernelLinOp.matvec/rmatvec/row_oracle/blk_oracle``, ``kernels/base.py:43-47,104-128``).
* ``torch.ops.rlaopt_b200.kernel_matmat`` — the same as a ``torch.library`` op, the
  registration pattern of the reference's own ops (``rlaopt/csrc/cpp/csc_matmat.cpp:83-87``).
"""
from __future__ import annotations

import os
from typing import Optional, Union

import torch

from . import _lib
from ._lib import LAYOUT_SIMT, LAYOUT_TC

KERNEL_IDS = {"rbf": 0, "laplace": 1, "matern12": 2, "matern32": 3, "matern52": 4}
_KERNEL_NAMES = {v: k for k, v in KERNEL_IDS.items()}

# Count of CUDA kernels this package has launched (bench.py reports it as gpu_launches).
LAUNCH_COUNT = 0

_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


def kernel_id(kernel: Union[str, int]) -> int:
    if isinstance(kernel, int):
        if kernel not in _KERNEL_NAMES:
            raise ValueError(f"unknown kernel id {kernel}")
        return kernel
    try:
        return KERNEL_IDS[kernel.lower()]
    except KeyError:
        raise ValueError(f"unknown kernel {kernel!r}; expected one of {sorted(KERNEL_IDS)}") from None


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"rlaopt_b200: {name} is on {t.device}; the kernel-matmat path runs only on CUDA devices "
            "(hand-written sm_100a kernels, no CPU fallback)"
        )


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class PackedPoints:
    """A point set in the fused kernels' streaming layout (device buffer + shape)."""

    __slots__ = ("buf", "n", "d", "dtype", "layout", "device")

    def __init__(self, buf: torch.Tensor, n: int, d: int, dtype: torch.dtype, layout: int):
        self.buf, self.n, self.d, self.dtype, self.layout = buf, n, d, dtype, layout
        self.device = buf.device


def choose_layout(kid: int, dtype: torch.dtype, d: int, k: int) -> int:
    """Kernel-path selection by shape (one backend, two kernels)."""
    forced = os.environ.get("RLAOPT_B200_LAYOUT", "").lower()
    if forced == "simt":
        return LAYOUT_SIMT
    elem = 4 if dtype == torch.float32 else 8
    tc_ok = bool(_lib.load().rlaopt_b200_layout_supported(kid, elem, d, k, LAYOUT_TC))
    # Matern-1/2 (exp(-r), not smooth at r = 0) also runs on the tensor-core path: its epilogue recomputes
    # (near-)coincident pairs from direct differences, where the GEMM-form distance has no relative accuracy.
    if forced == "tc":
        if not tc_ok:
            raise RuntimeError(f"RLAOPT_B200_LAYOUT=tc but kernel={_KERNEL_NAMES[kid]} dtype={dtype} d={d} k={k} unsupported")
        return LAYOUT_TC
    return LAYOUT_TC if tc_ok else LAYOUT_SIMT


def _inv_lengthscale(lengthscale, d: int, dtype: torch.dtype, device: torch.device):
    """(scalar, vector-or-None) of 1/lengthscale; vector for the per-feature (ARD) form."""
    if isinstance(lengthscale, torch.Tensor):
        if lengthscale.ndim == 0:
            return 1.0 / float(lengthscale), None
        if lengthscale.ndim != 1 or lengthscale.shape[0] != d:
            raise ValueError(f"lengthscale tensor must have shape ({d},), got {tuple(lengthscale.shape)}")
        vec = (1.0 / lengthscale.to(device=device, dtype=torch.float64)).to(dtype).contiguous()
        return 1.0, vec
    return 1.0 / float(lengthscale), None


def pack_points(
    X: torch.Tensor,
    lengthscale,
    idx: Optional[torch.Tensor] = None,
    layout: int = LAYOUT_SIMT,
) -> PackedPoints:
    """Gather ``X[idx]``, divide by the lengthscale, and lay out for the fused kernels."""
    global LAUNCH_COUNT
    _require_cuda(X, "X")
    if X.ndim != 2:
        raise ValueError(f"X must be 2-D, got {X.ndim}-D")
    if X.dtype not in _SUFFIX:
        raise ValueError(f"X dtype must be float32 or float64, got {X.dtype}")
    if X.stride(1) != 1 and X.shape[1] > 1:
        X = X.contiguous()
    lib = _lib.load()
    n_src, d = X.shape
    if idx is not None:
        idx = idx.to(device=X.device, dtype=torch.int64).contiguous()
        if idx.ndim != 1:
            raise ValueError("index tensor must be 1-D")
        n = idx.shape[0]
    else:
        n = n_src
    inv, inv_vec = _inv_lengthscale(lengthscale, d, X.dtype, X.device)
    elem = X.element_size()
    nbytes = lib.rlaopt_b200_packed_bytes(n, d, elem, layout)
    with torch.cuda.device(X.device):
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=X.device)
        fn = getattr(lib, f"rlaopt_b200_pack_points_{_SUFFIX[X.dtype]}")
        ldx = X.stride(0) if X.shape[0] > 1 else max(d, 1)
        rc = fn(_ptr(X), n, d, ldx, _ptr(idx), inv, _ptr(inv_vec), layout, _ptr(buf), _stream(X.device))
        _lib.check(rc, "pack_points")
    LAUNCH_COUNT += 2 if layout == LAYOUT_TC else 1  # TC: abs-max + split/pack kernels
    return PackedPoints(buf, n, d, X.dtype, layout)


def matmat_packed(
    rows: PackedPoints,
    cols: PackedPoints,
    V: torch.Tensor,
    kernel: Union[str, int],
    const_scaling: float = 1.0,
) -> torch.Tensor:
    """``c * K(rows, cols) @ V`` on packed operands; V is (m,) or (m, k), result matches."""
    global LAUNCH_COUNT
    kid = kernel_id(kernel)
    _require_cuda(V, "V")
    if rows.layout != cols.layout or rows.dtype != cols.dtype or rows.d != cols.d:
        raise ValueError("packed operands disagree in layout / dtype / feature count")
    if V.device != rows.device or cols.device != rows.device:
        raise ValueError("operands and V must be on the same device")
    if V.dtype != rows.dtype:
        raise ValueError(f"V has dtype {V.dtype}, operator has dtype {rows.dtype}")
    if V.ndim not in (1, 2):
        raise ValueError(f"x must be a 1D or 2D tensor. Received {V.ndim}D tensor.")
    if V.shape[0] != cols.n:
        raise ValueError(f"dimension mismatch: operator has {cols.n} columns, V has {V.shape[0]} rows")
    vec = V.ndim == 1
    Vm = V.unsqueeze(1) if vec else V
    if Vm.shape[1] > 1 and Vm.stride(1) != 1:
        Vm = Vm.contiguous()
    if Vm.shape[0] > 1 and Vm.stride(0) < Vm.shape[1]:
        Vm = Vm.contiguous()
    k = Vm.shape[1]
    lib = _lib.load()
    elem = V.element_size()
    with torch.cuda.device(V.device):
        Y = torch.empty((rows.n, k), dtype=V.dtype, device=V.device)
        ws_bytes = lib.rlaopt_b200_matmat_workspace_bytes(rows.n, cols.n, rows.d, k, elem, rows.layout)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=V.device) if ws_bytes else None
        fn = getattr(lib, f"rlaopt_b200_matmat_packed_{_SUFFIX[V.dtype]}")
        ldv = Vm.stride(0) if Vm.shape[0] > 1 else k
        rc = fn(
            _ptr(rows.buf), rows.n, _ptr(cols.buf), cols.n, rows.d, _ptr(Vm), k, ldv, _ptr(Y), k, kid,
            float(const_scaling), rows.layout, _ptr(ws), ws_bytes, _stream(V.device),
        )
        _lib.check(rc, "matmat_packed")
    if rows.layout == LAYOUT_TC:  # V split/pack + fused kernel (+ split reduce)
        v_bytes = -(-cols.n // 64) * (-(-k // 16) * 16) * 512
        LAUNCH_COUNT += 3 if ws_bytes > 2 * v_bytes else 2
    else:
        LAUNCH_COUNT += 2 if ws_bytes else 1
    return Y[:, 0] if vec else Y


def kernel_matmat(
    A1: torch.Tensor,
    A2: torch.Tensor,
    V: torch.Tensor,
    kernel: Union[str, int],
    lengthscale,
    const_scaling: float = 1.0,
    transpose: bool = False,
    row_idx: Optional[torch.Tensor] = None,
    col_idx: Optional[torch.Tensor] = None,
    layout: Optional[int] = None,
) -> torch.Tensor:
    """``c * K(A1[row_idx], A2[col_idx]) @ V`` (``transpose``: ``K^T @ V``), fused on the GPU."""
    kid = kernel_id(kernel)
    for name, t in (("A1", A1), ("A2", A2), ("V", V)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} is of type {type(t).__name__}, but expected type torch.Tensor")
        _require_cuda(t, name)
    if A1.ndim != 2 or A2.ndim != 2:
        raise ValueError("A1 and A2 must be 2D tensors")
    if A1.shape[1] != A2.shape[1]:
        raise ValueError(f"A1 and A2 must have the same number of features, got {A1.shape[1]} and {A2.shape[1]}")
    if A1.device != A2.device or A1.device != V.device:
        raise ValueError("A1, A2 and V must be on the same device.")
    if A1.dtype != A2.dtype or A1.dtype != V.dtype:
        raise ValueError("A1, A2 and V must have the same dtype.")
    k = 1 if V.ndim == 1 else V.shape[-1]
    if layout is None:
        layout = choose_layout(kid, A1.dtype, A1.shape[1], k)
    P1 = pack_points(A1, lengthscale, row_idx, layout)
    P2 = pack_points(A2, lengthscale, col_idx, layout)
    rows, cols = (P2, P1) if transpose else (P1, P2)
    return matmat_packed(rows, cols, V, kid, const_scaling)


# ---------------------------------------------------------------------------
# torch.library registration (same pattern as the reference's csc_matmat op).
# ---------------------------------------------------------------------------
_TORCH_LIB = torch.library.Library("rlaopt_b200", "FRAGMENT")
_TORCH_LIB.define(
    "kernel_matmat(Tensor A1, Tensor A2, Tensor V, int kernel_id, float lengthscale, Tensor? lengthscale_vec, "
    "float const_scaling, bool transpose=False, Tensor? row_idx=None, Tensor? col_idx=None) -> Tensor"
)


def _kernel_matmat_cuda(A1, A2, V, kernel_id_, lengthscale, lengthscale_vec, const_scaling, transpose=False,
                        row_idx=None, col_idx=None):
    ls = lengthscale_vec if lengthscale_vec is not None else lengthscale
    return kernel_matmat(A1, A2, V, kernel_id_, ls, const_scaling, transpose, row_idx, col_idx)


def _kernel_matmat_cpu(*args, **kwargs):
    raise RuntimeError(
        "rlaopt_b200::kernel_matmat has no CPU implementation: the kernel-matmat path is CUDA-only (sm_100a)"
    )


_TORCH_LIB.impl("kernel_matmat", _kernel_matmat_cuda, "CUDA")
_TORCH_LIB.impl("kernel_matmat", _kernel_matmat_cpu, "CPU")
