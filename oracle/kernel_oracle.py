"""CPU oracle for the implicit kernel-matrix matmat  Y = c * K(A1, A2) @ V.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rlaopt_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or as the timed CPU baseline, never as the product.

What it restates
----------------
The reference delegates the arithmetic of this path to PyKeOps ``LazyTensor``
(third-party: ``pykeops>=2.2.0`` in ``pyproject.toml:24``, dev pin
``pykeops==2.3`` / ``keopscore==2.3`` in ``requirements-dev.txt:48,18``; absent
from ``/root/reference`` and not installable here).  KeOps evaluates the
*symbolic formula* the reference writes down, so the oracle restates those
formulas dense and chunked in torch on the CPU:

* scaled difference ``(x_i - y_j) / lengthscale`` ........ ``rlaopt/kernels/standard.py:31-35``
* Matern distance ``sqrt(sum(u**2))`` ...................... ``rlaopt/kernels/standard.py:38-43``
* RBF ``exp(-sum(u**2) / 2)`` .............................. ``rlaopt/kernels/standard.py:46-52``
* Laplace ``exp(-sum(|u|))`` ............................... ``rlaopt/kernels/standard.py:55-61``
* Matern-1/2, -3/2, -5/2 ................................... ``rlaopt/kernels/standard.py:64-85``
* ``const_scaling`` multiplies the product, skipped at 1.0 . ``rlaopt/linops/mixins.py:26-29,60-72``
* forward / transpose reductions ``K @ x``, ``K.T @ x`` .... ``rlaopt/kernels/base.py:43-47``
* row / block oracles ``K(A1[blk], A2)``, ``K(A1[blk], A2[blk])`` ``rlaopt/kernels/base.py:88-128``

Pinning
-------
The reference holds no stored vectors for this path; its own tests pin the
operator against a closed-form double loop (``tests/kernels/utils.py:4-60``).
``oracle/gen_golden.py`` imports exactly that file from ``/root/reference`` and
stores its outputs on seeded inputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against them.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Union

import torch

KERNEL_NAMES = ("rbf", "laplace", "matern12", "matern32", "matern52")
KERNEL_IDS = {name: i for i, name in enumerate(KERNEL_NAMES)}

# rlaopt/kernels/standard.py:27-28
_SQRT3 = 3**0.5
_SQRT5 = 5**0.5

Lengthscale = Union[float, torch.Tensor]


def _as_kernel_name(kernel: Union[str, int]) -> str:
    if isinstance(kernel, int):
        return KERNEL_NAMES[kernel]
    name = kernel.lower()
    if name not in KERNEL_IDS:
        raise ValueError(f"unknown kernel {kernel!r}")
    return name


def _pointwise(name: str, u: torch.Tensor) -> torch.Tensor:
    """Kernel value from the scaled differences ``u`` of shape (r, m, d).

    Follows the exact operation order of ``rlaopt/kernels/standard.py:46-85``.
    """
    if name == "rbf":
        D = (u**2).sum(dim=2)  # standard.py:50-51
        return (-D / 2).exp()  # standard.py:52
    if name == "laplace":
        D = u.abs().sum(dim=2)  # standard.py:59-60
        return (-D).exp()  # standard.py:61
    D = (u**2).sum(dim=2).sqrt()  # standard.py:43
    if name == "matern12":
        return (-D).exp()  # standard.py:69
    if name == "matern32":
        return (1 + _SQRT3 * D) * (-_SQRT3 * D).exp()  # standard.py:77
    if name == "matern52":
        return (1 + _SQRT5 * D + 5 / 3 * D**2) * (-_SQRT5 * D).exp()  # standard.py:85
    raise ValueError(name)


def kernel_block(
    X: torch.Tensor, Y: torch.Tensor, kernel: Union[str, int], lengthscale: Lengthscale
) -> torch.Tensor:
    """Dense K(X, Y) (unscaled) in the dtype of X, direct-difference form.

    ``(X[:, None, :] - Y[None, :, :]) / lengthscale`` is the LazyTensor
    expression of ``rlaopt/kernels/standard.py:31-35`` evaluated eagerly.
    """
    name = _as_kernel_name(kernel)
    if isinstance(lengthscale, torch.Tensor):
        lengthscale = lengthscale.to(device=X.device, dtype=X.dtype)
    u = (X[:, None, :] - Y[None, :, :]) / lengthscale
    return _pointwise(name, u)


def kernel_matrix(
    A1: torch.Tensor,
    A2: torch.Tensor,
    kernel: Union[str, int],
    lengthscale: Lengthscale,
    const_scaling: float = 1.0,
    dtype: Optional[torch.dtype] = None,
) -> torch.Tensor:
    """Dense ``c * K(A1, A2)``; only for small problems."""
    if dtype is not None:
        A1, A2 = A1.to(dtype), A2.to(dtype)
    K = kernel_block(A1, A2, kernel, lengthscale)
    return K if const_scaling == 1.0 else const_scaling * K


def _rows_per_chunk(m: int, d: int, itemsize: int, budget_bytes: int = 256 << 20) -> int:
    return max(1, min(4096, budget_bytes // max(1, m * d * itemsize)))


def kernel_matmat(
    A1: torch.Tensor,
    A2: torch.Tensor,
    V: torch.Tensor,
    kernel: Union[str, int],
    lengthscale: Lengthscale,
    const_scaling: float = 1.0,
    transpose: bool = False,
    row_idx: Optional[torch.Tensor] = None,
    col_idx: Optional[torch.Tensor] = None,
    dtype: Optional[torch.dtype] = None,
    chunk: Optional[int] = None,
) -> torch.Tensor:
    """``c * K(A1[row_idx], A2[col_idx]) @ V`` (or ``K.T @ V``), chunked over rows.

    * forward / transpose ........ ``rlaopt/kernels/base.py:43-47``
    * ``row_idx`` / ``col_idx`` ... ``rlaopt/kernels/base.py:88-102`` (oracles)
    * post-scaling ............... ``rlaopt/linops/mixins.py:26-29``; skipped when
      ``const_scaling == 1.0`` (``mixins.py:60-61``)

    ``dtype=torch.float64`` gives the ground truth used for the 1e-5 claim;
    ``dtype=None`` computes in the inputs' dtype (the fp32 "numerical twin" of
    the KeOps reduction, which also uses direct differences).
    V may be 1-D (matvec) or 2-D (matmat); the result has the same rank.
    """
    name = _as_kernel_name(kernel)
    if dtype is not None:
        A1, A2, V = A1.to(dtype), A2.to(dtype), V.to(dtype)
        if isinstance(lengthscale, torch.Tensor):
            lengthscale = lengthscale.to(dtype)
    if row_idx is not None:
        A1 = A1[row_idx]
    if col_idx is not None:
        A2 = A2[col_idx]
    rows, cols = (A2, A1) if transpose else (A1, A2)  # K.T = K(A2, A1) for these kernels
    vec = V.ndim == 1
    Vm = V[:, None] if vec else V
    if Vm.shape[0] != cols.shape[0]:
        raise ValueError(f"V has {Vm.shape[0]} rows, operator has {cols.shape[0]} columns")
    n, d = rows.shape
    step = chunk or _rows_per_chunk(cols.shape[0], d, rows.element_size())
    out = torch.empty(n, Vm.shape[1], dtype=rows.dtype)
    for s in range(0, n, step):
        out[s : s + step] = kernel_block(rows[s : s + step], cols, name, lengthscale) @ Vm
    if const_scaling != 1.0:
        out = const_scaling * out
    return out[:, 0] if vec else out


def kernel_matmat_gemm_form(
    A1: torch.Tensor,
    A2: torch.Tensor,
    V: torch.Tensor,
    kernel: Union[str, int],
    lengthscale: Lengthscale,
    const_scaling: float = 1.0,
    chunk: int = 2048,
    row_idx: Optional[torch.Tensor] = None,
    dtype: Optional[torch.dtype] = None,
) -> torch.Tensor:
    """Fastest honest CPU path for the L2 kernels: ``|x|^2 + |y|^2 - 2 x.y`` via GEMM.

    Same formulas (``rlaopt/kernels/standard.py:38-85``) with the squared distance
    expanded; used only as the timed ``cpu_baseline`` / ``--impl reference`` leg
    of ``bench.py``.  Laplace has no GEMM form and falls back to
    :func:`kernel_matmat`.
    """
    name = _as_kernel_name(kernel)
    if name == "laplace":
        return kernel_matmat(A1, A2, V, name, lengthscale, const_scaling, row_idx=row_idx, dtype=dtype)
    if row_idx is not None:
        A1 = A1[row_idx]
    if dtype is not None:  # fp64: ground truth for large sampled-row checks (cancellation ~1e-15)
        A1, A2, V = A1.to(dtype), A2.to(dtype), V.to(dtype)
    if isinstance(lengthscale, torch.Tensor):
        lengthscale = lengthscale.to(A1.dtype)
    X, Y = A1 / lengthscale, A2 / lengthscale
    vec = V.ndim == 1
    Vm = V[:, None] if vec else V
    yn = (Y * Y).sum(dim=1)[None, :]
    Yt = Y.T.contiguous()
    out = torch.empty(X.shape[0], Vm.shape[1], dtype=X.dtype)
    for s in range(0, X.shape[0], chunk):
        Xc = X[s : s + chunk]
        D = torch.addmm(yn + (Xc * Xc).sum(dim=1)[:, None], Xc, Yt, alpha=-2.0).clamp_(min=0.0)
        if name == "rbf":
            P = D.mul_(-0.5).exp_()
        else:
            r = D.sqrt_()
            if name == "matern12":
                P = r.neg_().exp_()
            elif name == "matern32":
                P = (1 + _SQRT3 * r) * (-_SQRT3 * r).exp()
            else:
                P = (1 + _SQRT5 * r + 5 / 3 * r * r) * (-_SQRT5 * r).exp()
        out[s : s + chunk] = P @ Vm
    if const_scaling != 1.0:
        out = const_scaling * out
    return out[:, 0] if vec else out


def kernel_entry_python(x: Sequence[float], y: Sequence[float], kernel, lengthscale) -> float:
    """One K_ij with python floats (double precision) — for tiny known-answer checks.

    ``lengthscale`` is a float or a sequence of length d.
    """
    name = _as_kernel_name(kernel)
    d = len(x)
    ls = [float(lengthscale)] * d if not hasattr(lengthscale, "__len__") else [float(v) for v in lengthscale]
    u = [(float(x[t]) - float(y[t])) / ls[t] for t in range(d)]
    if name == "laplace":
        return math.exp(-sum(abs(v) for v in u))
    sq = sum(v * v for v in u)
    if name == "rbf":
        return math.exp(-sq / 2)
    r = math.sqrt(sq)
    if name == "matern12":
        return math.exp(-r)
    if name == "matern32":
        return (1 + _SQRT3 * r) * math.exp(-_SQRT3 * r)
    return (1 + _SQRT5 * r + 5 / 3 * r * r) * math.exp(-_SQRT5 * r)


def row_chunks(n: int, world: int) -> list[torch.Tensor]:
    """Row partition of the distributed operator: ``torch.chunk(arange(n), g)``.

    ``rlaopt/kernels/base.py:297-302`` (A1_row_chunks / A2_row_chunks) and ``:462``
    (block chunks for ``blk_oracle``).
    """
    return list(torch.chunk(torch.arange(n), world, dim=0))


def rel_fro_error(Y: torch.Tensor, Y_ref: torch.Tensor) -> float:
    """``|Y - Y_ref|_F / |Y_ref|_F`` in float64 (the parity metric)."""
    Yd, Rd = Y.detach().double().cpu(), Y_ref.detach().double().cpu()
    den = torch.linalg.norm(Rd).item()
    num = torch.linalg.norm(Yd - Rd).item()
    return num / den if den > 0 else num
