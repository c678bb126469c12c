"""ctypes binding of the C ABI declared in ``include/rlaopt_b200.h``.

The shared library ``rlaopt_b200/csrc/librlaopt_b200.so`` is built in-tree by
``python -m rlaopt_b200.csrc.build`` (``__graft_entry__.build()``).  There is no
CPU fallback: if the library is missing, or no CUDA device is usable, every
compute entry raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# RLAOPT_B200_LIB selects another build of the same ABI (e.g. the -DKMM_TC_PROFILE diagnostic build)
LIB_PATH = os.environ.get("RLAOPT_B200_LIB") or os.path.join(_HERE, "csrc", "librlaopt_b200.so")

LAYOUT_SIMT = 0
LAYOUT_TC = 1


def _epilogue_struct(real):
    """ctypes mirror of ``rlaopt_b200_epilogue_f32 / _f64`` (include/rlaopt_b200.h)."""

    class Epilogue(ctypes.Structure):
        _fields_ = [
            ("alpha", real), ("beta", real),
            ("addend", c_void_p), ("ld_addend", c_int64), ("addend_rows", c_int64), ("addend_idx", c_void_p),
            ("gamma", real),
            ("rhs", c_void_p), ("ld_rhs", c_int64), ("rhs_rows", c_int64), ("rhs_idx", c_void_p),
            ("gram_lhs", c_void_p), ("ld_gram_lhs", c_int64), ("gram_cols", c_int64), ("gram_out", c_void_p),
            ("sqnorm_out", c_void_p),
        ]

    return Epilogue


EpilogueF32 = _epilogue_struct(c_float)
EpilogueF64 = _epilogue_struct(c_double)

_lib = None

# name -> (restype, argtypes); mirrors include/rlaopt_b200.h one to one
_f32p, _f64p, _i64p = c_void_p, c_void_p, c_void_p
PROTOTYPES = {
    "rlaopt_b200_abi_version": (c_int, []),
    "rlaopt_b200_last_error": (c_char_p, []),
    "rlaopt_b200_device_sm_count": (c_int, []),
    "rlaopt_b200_layout_supported": (c_int, [c_int, c_int, c_int64, c_int64, c_int]),
    "rlaopt_b200_packed_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "rlaopt_b200_pack_points_f32": (
        c_int,
        [_f32p, c_int64, c_int64, c_int64, c_int64, _i64p, c_float, _f32p, _f32p, c_int, c_void_p, c_void_p],
    ),
    "rlaopt_b200_pack_points_f64": (
        c_int,
        [_f64p, c_int64, c_int64, c_int64, c_int64, _i64p, c_double, _f64p, _f64p, c_int, c_void_p, c_void_p],
    ),
    "rlaopt_b200_column_mean_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "rlaopt_b200_column_mean_f32": (
        c_int,
        [_f32p, c_int64, c_int64, c_int64, c_int64, _i64p, _f32p, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_packed_stats_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rlaopt_b200_matmat_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int, c_int]),
    "rlaopt_b200_matmat_packed_f32": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, _f32p, c_int64, c_int64, _f32p, c_int64, c_int, c_float,
         c_int, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_matmat_packed_f64": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, _f64p, c_int64, c_int64, _f64p, c_int64, c_int, c_double,
         c_int, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_matmat_fused_workspace_bytes": (
        c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_int]),
    "rlaopt_b200_matmat_packed_fused_f32": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, _f32p, c_int64, c_int64, _f32p, c_int64, c_int, c_float,
         c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_matmat_packed_fused_f64": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, _f64p, c_int64, c_int64, _f64p, c_int64, c_int, c_double,
         c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_kernel_matmat_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int, c_int]),
    "rlaopt_b200_kernel_matmat_f32": (
        c_int,
        [_f32p, c_int64, c_int64, _f32p, c_int64, c_int64, c_int64, _f32p, c_int64, c_int64, _f32p, c_int64, c_int,
         c_float, _f32p, c_float, c_int, _i64p, c_int64, _i64p, c_int64, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_kernel_matmat_f64": (
        c_int,
        [_f64p, c_int64, c_int64, _f64p, c_int64, c_int64, c_int64, _f64p, c_int64, c_int64, _f64p, c_int64, c_int,
         c_double, _f64p, c_double, c_int, _i64p, c_int64, _i64p, c_int64, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "rlaopt_b200_kernel_matmat_host_f32": (
        c_int,
        [_f32p, c_int64, _f32p, c_int64, c_int64, _f32p, c_int64, _f32p, c_int, c_float, c_float, c_int, c_int],
    ),
}


def load():
    """Load (once) and return the ctypes handle; raise loudly if the build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"rlaopt_b200: CUDA extension not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `python -m rlaopt_b200.csrc.build`. "
            "There is no CPU fallback for the kernel-matmat path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """Turn a non-zero ABI return code into RuntimeError (reference ops raise via TORCH_CHECK)."""
    if rc != 0:
        msg = load().rlaopt_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"rlaopt_b200.{what} failed (code {rc}): {msg}")
