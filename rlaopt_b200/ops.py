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

import ctypes
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
    """A point set in the fused kernels' streaming layout (device buffer + shape).

    ``center`` is the shift that was subtracted before the division by the lengthscale (tensor-core layout
    only); two packs may be multiplied only if they were packed with the same shift.
    """

    __slots__ = ("buf", "n", "d", "dtype", "layout", "device", "center", "_stats")

    def __init__(self, buf: torch.Tensor, n: int, d: int, dtype: torch.dtype, layout: int, center=None):
        self.buf, self.n, self.d, self.dtype, self.layout = buf, n, d, dtype, layout
        self.device = buf.device
        self.center = center
        self._stats = None

    def stats(self) -> tuple[float, int]:
        """``(max_i |(x_i - center) / lengthscale|^2, number of out-of-range gather indices)`` of a tensor-core
        pack, read back once (one 8-byte copy and a stream synchronisation) and memoised."""
        if self._stats is None:
            if self.layout != LAYOUT_TC:
                raise RuntimeError("only tensor-core packs carry statistics")
            if self.n == 0:
                self._stats = (0.0, 0)
                return self._stats
            mx, bad = ctypes.c_float(0.0), ctypes.c_int64(0)
            with torch.cuda.device(self.device):
                rc = _lib.load().rlaopt_b200_packed_stats_host(
                    _ptr(self.buf), self.layout, ctypes.addressof(mx), ctypes.addressof(bad), _stream(self.device))
            _lib.check(rc, "packed_stats")
            self._stats = (float(mx.value), int(bad.value))
        return self._stats

    @property
    def max_sqnorm(self) -> float:
        return self.stats()[0]


# Accuracy guard of the tensor-core path.  Its GEMM-form distance carries an absolute error
# ~ TC_EPS_D (|x|^2 + |y|^2); the relative error this leaves in a kernel value is that times |d ln f / d D|, at most
# 1/2 (RBF), 3/2 (Matern-3/2), 5/6 (Matern-5/2); Matern-1/2 recomputes near pairs exactly and is bounded by 1e-5 / N
# beyond them.  While the bound stays below the 1e-5 parity bar the packs run on tcgen05; otherwise the operator
# falls back to the direct-difference CUDA-core kernel (exact under translation, like the reference's KeOps formula).
TC_EPS_D = 3.0e-7
TC_PARITY_TOL = 1.0e-5
TC_SENSITIVITY = {0: 0.5, 2: 0.5, 3: 1.5, 4: 5.0 / 6.0}


def tc_norm_budget(kid: int) -> float:
    """Largest ``max|x|^2 + max|y|^2`` (centred, lengthscale-scaled) the tensor-core path accepts for kernel ``kid``."""
    return TC_PARITY_TOL / (TC_EPS_D * TC_SENSITIVITY[kid])


def tc_accuracy_ok(kid: int, rows_sqnorm: float, cols_sqnorm: float) -> bool:
    if os.environ.get("RLAOPT_B200_LAYOUT", "").lower() == "tc":
        return True
    return rows_sqnorm + cols_sqnorm <= tc_norm_budget(kid)


def choose_layout(kid: int, dtype: torch.dtype, d: int, k: int) -> int:
    """Kernel-path selection by shape (one backend, two kernels); the data-dependent half of the decision is
    :func:`tc_accuracy_ok`, evaluated once per operator from the pack statistics."""
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


def _check_index(idx: torch.Tensor, n_src: int, device: torch.device) -> torch.Tensor:
    """Validate a gather list like ``A1[blk]`` does (IndexError on the host; a device-side assert for CUDA
    indices) and return it as a contiguous int64 tensor on ``device``.  Negative entries wrap in the kernels."""
    if idx.ndim != 1:
        raise ValueError("index tensor must be 1-D")
    if idx.dtype not in (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8):
        raise IndexError(f"index tensor must be an integer tensor, got {idx.dtype}")
    if idx.numel() > 0:
        if idx.device.type == "cpu":
            lo, hi = int(idx.min()), int(idx.max())
            if lo < -n_src or hi >= n_src:
                bad = lo if lo < -n_src else hi
                raise IndexError(f"index {bad} is out of bounds for dimension 0 with size {n_src}")
        else:
            torch._assert_async(((idx >= -n_src) & (idx < n_src)).all(),
                                f"rlaopt_b200: gather index out of bounds for dimension 0 with size {n_src}")
    return idx.to(device=device, dtype=torch.int64).contiguous()


def column_mean(X: torch.Tensor, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-feature mean of ``X[idx]`` (fp64-accumulated on the device, bit-reproducible): the common shift of the
    tensor-core packs of one operator."""
    global LAUNCH_COUNT
    _require_cuda(X, "X")
    if X.dtype != torch.float32 or X.ndim != 2:
        raise ValueError("column_mean expects a 2-D float32 tensor")
    if X.stride(1) != 1 and X.shape[1] > 1:
        X = X.contiguous()
    lib = _lib.load()
    n_src, d = X.shape
    if idx is not None:
        idx = _check_index(idx, n_src, X.device)
    n = n_src if idx is None else idx.shape[0]
    with torch.cuda.device(X.device):
        center = torch.zeros(d, dtype=torch.float32, device=X.device)
        if n == 0:
            return center
        ws_bytes = lib.rlaopt_b200_column_mean_workspace_bytes(n, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
        ldx = X.stride(0) if X.shape[0] > 1 else max(d, 1)
        rc = lib.rlaopt_b200_column_mean_f32(_ptr(X), n, n_src, d, ldx, _ptr(idx), _ptr(center), _ptr(ws), ws_bytes,
                                             _stream(X.device))
        _lib.check(rc, "column_mean")
    LAUNCH_COUNT += 2
    return center


def pack_points(
    X: torch.Tensor,
    lengthscale,
    idx: Optional[torch.Tensor] = None,
    layout: int = LAYOUT_SIMT,
    center: Optional[torch.Tensor] = None,
) -> PackedPoints:
    """Gather ``X[idx]``, subtract ``center`` (tensor-core layout), divide by the lengthscale, and lay out for
    the fused kernels."""
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
        idx = _check_index(idx, n_src, X.device)
        n = idx.shape[0]
    else:
        n = n_src
    if layout != LAYOUT_TC:
        center = None  # direct differences are exact under translation
    elif center is not None:
        if center.shape != (d,):
            raise ValueError(f"center must have shape ({d},), got {tuple(center.shape)}")
        center = center.to(device=X.device, dtype=X.dtype).contiguous()
    inv, inv_vec = _inv_lengthscale(lengthscale, d, X.dtype, X.device)
    elem = X.element_size()
    nbytes = lib.rlaopt_b200_packed_bytes(n, d, elem, layout)
    with torch.cuda.device(X.device):
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=X.device)
        fn = getattr(lib, f"rlaopt_b200_pack_points_{_SUFFIX[X.dtype]}")
        ldx = X.stride(0) if X.shape[0] > 1 else max(d, 1)
        rc = fn(_ptr(X), n, n_src, d, ldx, _ptr(idx), inv, _ptr(inv_vec), _ptr(center), layout, _ptr(buf),
                _stream(X.device))
        _lib.check(rc, "pack_points")
    LAUNCH_COUNT += 2 if layout == LAYOUT_TC else 1  # TC: abs-max + split/pack kernels
    return PackedPoints(buf, n, d, X.dtype, layout, center)


def _same_center(a: Optional[torch.Tensor], b: Optional[torch.Tensor]) -> bool:
    if a is None or b is None:
        return a is b
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape)


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
    if not _same_center(rows.center, cols.center):
        raise ValueError("packed operands were shifted by different centers: K(x - c1, y - c2) is not K(x, y)")
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


def _rows_operand(t: torch.Tensor, k: int, name: str, like: torch.Tensor) -> torch.Tensor:
    """(rows, k) operand of the fused output stage: 2-D view, unit column stride, on the product's device."""
    if t.device != like.device or t.dtype != like.dtype:
        raise ValueError(f"{name} must live on {like.device} with dtype {like.dtype}")
    t2 = t.unsqueeze(1) if t.ndim == 1 else t
    if t2.ndim != 2 or t2.shape[1] != k:
        raise ValueError(f"{name} must have {k} column(s), got shape {tuple(t.shape)}")
    if (t2.shape[1] > 1 and t2.stride(1) != 1) or (t2.shape[0] > 1 and t2.stride(0) < t2.shape[1]):
        t2 = t2.contiguous()
    return t2


def matmat_packed_fused(
    rows: PackedPoints,
    cols: PackedPoints,
    V: torch.Tensor,
    kernel: Union[str, int],
    const_scaling: float = 1.0,
    *,
    alpha: float = 1.0,
    addend: Optional[torch.Tensor] = None,
    beta: float = 0.0,
    addend_idx: Optional[torch.Tensor] = None,
    rhs: Optional[torch.Tensor] = None,
    gamma: float = 0.0,
    rhs_idx: Optional[torch.Tensor] = None,
    gram_with: Optional[torch.Tensor] = None,
    want_sqnorm: bool = False,
    store: bool = True,
):
    """``Y = alpha * c * K(rows, cols) @ V + beta * addend[addend_idx] + gamma * rhs[rhs_idx]`` with the Gram
    matrix ``gram_with^T Y`` and the squared column norms of ``Y`` taken in the same pass (C entry
    ``rlaopt_b200_matmat_packed_fused_*``).  Returns ``(Y or None, gram or None, sqnorm or None)``; ``store=False``
    skips writing ``Y`` (only the reductions are wanted -- no n x k result exists anywhere)."""
    global LAUNCH_COUNT
    kid = kernel_id(kernel)
    _require_cuda(V, "V")
    if rows.layout != cols.layout or rows.dtype != cols.dtype or rows.d != cols.d:
        raise ValueError("packed operands disagree in layout / dtype / feature count")
    if not _same_center(rows.center, cols.center):
        raise ValueError("packed operands were shifted by different centers: K(x - c1, y - c2) is not K(x, y)")
    if V.device != rows.device or cols.device != rows.device:
        raise ValueError("operands and V must be on the same device")
    if V.dtype != rows.dtype:
        raise ValueError(f"V has dtype {V.dtype}, operator has dtype {rows.dtype}")
    if V.ndim not in (1, 2):
        raise ValueError(f"x must be a 1D or 2D tensor. Received {V.ndim}D tensor.")
    if V.shape[0] != cols.n:
        raise ValueError(f"dimension mismatch: operator has {cols.n} columns, V has {V.shape[0]} rows")
    vec = V.ndim == 1
    k = 1 if vec else V.shape[1]
    Vm = _rows_operand(V, k, "V", V)
    n = rows.n
    lib = _lib.load()
    f32 = V.dtype == torch.float32
    epi = (_lib.EpilogueF32 if f32 else _lib.EpilogueF64)()
    keep = []  # tensors the launch reads: kept alive until it is enqueued
    epi.alpha, epi.beta, epi.gamma = float(alpha), float(beta), float(gamma)
    for name, t, idx in (("addend", addend, addend_idx), ("rhs", rhs, rhs_idx)):
        if t is None:
            continue
        t2 = _rows_operand(t, k, name, V)
        if idx is not None:
            idx = _check_index(idx, t2.shape[0], V.device)
            if idx.shape[0] != n:
                raise ValueError(f"{name}_idx must have one entry per output row ({n}), got {idx.shape[0]}")
        elif t2.shape[0] != n:
            raise ValueError(f"{name} must have {n} rows, got {t2.shape[0]}")
        keep += [t2, idx]
        setattr(epi, name, t2.data_ptr())
        setattr(epi, f"ld_{name}", t2.stride(0) if t2.shape[0] > 1 else k)
        setattr(epi, f"{name}_rows", t2.shape[0])
        setattr(epi, f"{name}_idx", _ptr(idx))
    gram = sqn = None
    gcols = 0
    with torch.cuda.device(V.device):
        if gram_with is not None:
            gcols = 1 if gram_with.ndim == 1 else gram_with.shape[1]
            L = _rows_operand(gram_with, gcols, "gram_with", V)
            if L.shape[0] != n:
                raise ValueError(f"gram_with must have {n} rows, got {L.shape[0]}")
            gram = torch.empty((gcols, k), dtype=V.dtype, device=V.device)
            keep.append(L)
            epi.gram_lhs, epi.ld_gram_lhs, epi.gram_cols = L.data_ptr(), (L.stride(0) if n > 1 else gcols), gcols
            epi.gram_out = gram.data_ptr()
        if want_sqnorm:
            sqn = torch.empty(k, dtype=V.dtype, device=V.device)
            epi.sqnorm_out = sqn.data_ptr()
        Y = torch.empty((n, k), dtype=V.dtype, device=V.device) if store else None
        elem = V.element_size()
        ws_bytes = lib.rlaopt_b200_matmat_fused_workspace_bytes(n, cols.n, rows.d, k, elem, rows.layout, gcols,
                                                                int(want_sqnorm))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=V.device)
        fn = getattr(lib, f"rlaopt_b200_matmat_packed_fused_{_SUFFIX[V.dtype]}")
        ldv = Vm.stride(0) if Vm.shape[0] > 1 else k
        rc = fn(_ptr(rows.buf), n, _ptr(cols.buf), cols.n, rows.d, _ptr(Vm), k, ldv, _ptr(Y), k, kid,
                float(const_scaling), rows.layout, ctypes.addressof(epi), _ptr(ws), ws_bytes, _stream(V.device))
        _lib.check(rc, "matmat_packed_fused")
    LAUNCH_COUNT += 3 + (1 if (gram is not None or want_sqnorm) else 0)  # V pack, fused kernel, output stage (+ reduce)
    del keep
    if Y is not None and vec:
        Y = Y[:, 0]
    return Y, gram, sqn


def fused_reductions_supported(k: int, gram_cols: int = 0) -> bool:
    """The Gram / column-norm reductions of the fused output stage cover k <= 64 and gram_cols <= 64."""
    return 1 <= k <= 64 and 0 <= gram_cols <= 64


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
    auto = layout is None
    if auto:
        layout = choose_layout(kid, A1.dtype, A1.shape[1], k)
    center = column_mean(A2, col_idx) if layout == LAYOUT_TC else None
    P1 = pack_points(A1, lengthscale, row_idx, layout, center)
    P2 = pack_points(A2, lengthscale, col_idx, layout, center)
    if auto and layout == LAYOUT_TC and not tc_accuracy_ok(kid, P1.max_sqnorm, P2.max_sqnorm):
        # norms too large for the GEMM-form distance at the parity bar: direct-difference kernel instead
        P1 = pack_points(A1, lengthscale, row_idx, LAYOUT_SIMT)
        P2 = pack_points(A2, lengthscale, col_idx, LAYOUT_SIMT)
    rows, cols = (P2, P1) if transpose else (P1, P2)
    return matmat_packed(rows, cols, V, kid, const_scaling)


# ---------------------------------------------------------------------------
# torch.library registration.
#
# * ``torch.ops.rlaopt.kernel_matmat`` -- C++ (``csrc/torch_op.cpp``: TORCH_LIBRARY_FRAGMENT(rlaopt, m) + CUDA / CPU
#   implementations in ``librlaopt_b200_torch.so``), the registration pattern of the reference's own ops
#   (``rlaopt/csrc/cpp/csc_matmat.cpp:83-87``); loaded by :func:`load_torch_op`.
# * ``torch.ops.rlaopt_b200.kernel_matmat`` -- the same schema registered from Python over this module's functions
#   (kept for callers of round 1).
# ---------------------------------------------------------------------------
TORCH_OP_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "librlaopt_b200_torch.so")
_torch_op_loaded = False


def load_torch_op() -> bool:
    """Load ``librlaopt_b200_torch.so`` (once); afterwards ``torch.ops.rlaopt.kernel_matmat`` exists.  Returns False
    when the library has not been built (``python -m rlaopt_b200.csrc.build``)."""
    global _torch_op_loaded
    if not _torch_op_loaded and os.path.exists(TORCH_OP_PATH):
        _lib.load()  # the op library links the C-ABI library
        torch.ops.load_library(TORCH_OP_PATH)
        _torch_op_loaded = True
    return _torch_op_loaded


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
