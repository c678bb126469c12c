"""Kernel linear operators: single-device and row-partitioned multi-GPU.

Interface of ``rlaopt/kernels/base.py``:

* ``_KernelLinOp`` (``base.py:23-128``) — ``K @ x``, ``K.T @ x``, ``row_oracle(blk)``,
  ``blk_oracle(blk)``, properties ``A1 / A2 / kernel_config``;
* ``_DistributedKernelLinOp`` (``base.py:247-520``) — the same operator with the rows
  of ``A1`` partitioned over a set of devices, ``A2`` replicated, plus distributed
  row / block oracles and ``shutdown()``.

Where the reference builds PyKeOps ``LazyTensor`` formulas and reduces them with
``K_lazy @ x``, these classes pack the point sets once into the layout of the
fused CUDA kernels (``rlaopt_b200.ops.pack_points``) and call
``rlaopt_b200.ops.matmat_packed``; ``const_scaling`` is applied in the kernel
epilogue instead of a second pass (``rlaopt/linops/mixins.py:26-29``).  The
reference's per-process LazyTensor caches (``base.py:19-20,183-244``) become
per-operator pack caches.  There is no CPU path: operators can be *constructed*
from CPU tensors (shape / validation logic is host code), applying them raises.
"""
from __future__ import annotations

from typing import Any, Optional

import torch

from rlaopt_b200 import ops
from rlaopt_b200.linops import DistributedTwoSidedLinOp, LinOp, ScaleMixin, TwoSidedLinOp
from rlaopt_b200.linops.distributed import _DistributedLinOp
from rlaopt_b200.utils import _is_set, _is_torch_tensor

from .configs import KernelConfig, _is_kernel_config


def _same_storage(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (
        a.device == b.device
        and a.shape == b.shape
        and a.stride() == b.stride()
        and a.dtype == b.dtype
        and a.data_ptr() == b.data_ptr()
    )


def _n_cols(x: torch.Tensor) -> int:
    return 1 if x.ndim == 1 else x.shape[1]


def _ls_state(lengthscale):
    """Identity + in-place version of a lengthscale (tensor) or its value (float)."""
    if isinstance(lengthscale, torch.Tensor):
        return (id(lengthscale), lengthscale._version)
    return float(lengthscale)


class _PackCache:
    """Packed forms of (A1, A2) per kernel layout; ``A1 is A2`` shares one pack.

    Tensor-core packs are shifted by one common center, the column means of ``A2`` (all kernels are functions of
    ``x - y``, ``rlaopt/kernels/standard.py:31-43``, so K is unchanged).  The reference's LazyTensor reads the live
    tensors on every product; the packs are snapshots, so they are keyed on the in-place versions of ``A1``, ``A2``
    and the lengthscale and rebuilt when any of them was modified.
    """

    def __init__(self, A1: torch.Tensor, A2: torch.Tensor, kernel_config):
        self.A1, self.A2, self.cfg = A1, A2, kernel_config
        self.shared = _same_storage(A1, A2)
        self._rows: dict[int, object] = {}
        self._cols: dict[int, object] = {}
        self._center: Optional[torch.Tensor] = None
        self._tc_ok: dict[int, bool] = {}
        self._state = self._live_state()

    @property
    def lengthscale(self):
        return self.cfg.lengthscale

    def _live_state(self):
        return (self.A1._version, self.A2._version, _ls_state(self.cfg.lengthscale))

    def validate(self) -> None:
        """Drop every pack if A1, A2 or the lengthscale changed since they were built."""
        state = self._live_state()
        if state != self._state:
            self.clear()
            self._state = state

    def center(self) -> Optional[torch.Tensor]:
        """Common shift of the tensor-core packs (fp32 only; ``None`` for fp64 operators)."""
        self.validate()
        if self._center is None and self.A2.dtype == torch.float32 and self.A2.is_cuda:
            self._center = ops.column_mean(self.A2)
        return self._center

    def _pack(self, A: torch.Tensor, layout: int):
        return ops.pack_points(A, self.lengthscale, None, layout, self.center() if layout == ops.LAYOUT_TC else None)

    def rows(self, layout: int):
        """Pack of A1 (built on first use; shared with the column pack when A1 is A2)."""
        self.validate()
        if layout not in self._rows:
            if self.shared and layout in self._cols:
                self._rows[layout] = self._cols[layout]
            else:
                self._rows[layout] = self._pack(self.A1, layout)
        return self._rows[layout]

    def cols(self, layout: int):
        """Pack of A2 alone -- all a row oracle needs (its rows are gathered per block)."""
        self.validate()
        if layout not in self._cols:
            if self.shared and layout in self._rows:
                self._cols[layout] = self._rows[layout]
            else:
                self._cols[layout] = self._pack(self.A2, layout)
        return self._cols[layout]

    def get(self, layout: int):
        return self.rows(layout), self.cols(layout)

    def known_sqnorm(self, which: str) -> Optional[float]:
        """Largest centred squared norm of A1 / A2 if its tensor-core pack exists (an upper bound for any row
        subset packed with the same center), else ``None``."""
        packs = self._rows if which == "rows" else self._cols
        other = self._cols if which == "rows" else self._rows
        P = packs.get(ops.LAYOUT_TC) or (other.get(ops.LAYOUT_TC) if self.shared else None)
        return None if P is None else P.max_sqnorm

    def tc_ok(self, kid: int) -> bool:
        """Data-dependent half of the layout choice for the full operator (one 8-byte read-back per operator)."""
        self.validate()
        if kid not in self._tc_ok:
            P1, P2 = self.get(ops.LAYOUT_TC)
            ok = ops.tc_accuracy_ok(kid, P1.max_sqnorm, P2.max_sqnorm)
            self._tc_ok[kid] = ok
            if not ok:  # the tensor-core packs will not be used: release them
                self._rows.pop(ops.LAYOUT_TC, None)
                self._cols.pop(ops.LAYOUT_TC, None)
        return self._tc_ok[kid]

    def clear(self) -> None:
        self._rows.clear()
        self._cols.clear()
        self._center = None
        self._tc_ok.clear()


class _KernelLinOp(TwoSidedLinOp, ScaleMixin):
    """``const_scaling * K(A1, A2)`` as a two-sided linear operator on one GPU."""

    def __init__(self, A1: torch.Tensor, A2: torch.Tensor, kernel_config: KernelConfig, _kernel_key: str):
        self._check_inputs(A1, A2, kernel_config)
        self._A1, self._A2 = A1, A2
        self._kernel_config = kernel_config
        self._kernel_key = _kernel_key
        self._kernel_id = ops.kernel_id(_kernel_key)
        self._initialize_scaling(getattr(kernel_config, "const_scaling", 1.0))
        self._cache = _PackCache(A1, A2, kernel_config)
        self._oracle_memo: dict[str, tuple] = {}
        super().__init__(
            device=A1.device,
            shape=torch.Size((A1.shape[0], A2.shape[0])),
            matvec=self._forward,
            rmatvec=self._adjoint,
            matmat=self._forward,  # the fused kernel handles any number of columns
            rmatmat=self._adjoint,
            dtype=A1.dtype,
        )

    # -- properties ---------------------------------------------------------
    @property
    def A1(self) -> torch.Tensor:
        return self._A1

    @property
    def A2(self) -> torch.Tensor:
        return self._A2

    @property
    def kernel_config(self) -> KernelConfig:
        return self._kernel_config

    def _check_inputs(self, A1: Any, A2: Any, kernel_config: Any) -> None:
        _is_torch_tensor(A1, "A1")
        _is_torch_tensor(A2, "A2")
        if A1.ndim != 2:
            raise ValueError(f"A1 must be a 2D tensor, got {A1.ndim}D tensor.")
        if A2.ndim != 2:
            raise ValueError(f"A2 must be a 2D tensor, got {A2.ndim}D tensor.")
        if A1.device != A2.device:
            raise ValueError("A1 and A2 must be on the same device.")
        if A1.dtype != A2.dtype:
            raise ValueError("A1 and A2 must have the same dtype.")
        if A1.shape[1] != A2.shape[1]:
            raise ValueError(
                f"A1 and A2 must have the same number of features, got {A1.shape[1]} and {A2.shape[1]}."
            )
        _is_kernel_config(kernel_config, "kernel_config")

    # -- products -----------------------------------------------------------
    def _layout_for(self, x: torch.Tensor) -> int:
        """Shape rule (``ops.choose_layout``) plus the accuracy guard of the tensor-core path: data whose centred,
        lengthscale-scaled norms exceed the budget of ``ops.tc_accuracy_ok`` runs on the direct-difference kernel."""
        layout = ops.choose_layout(self._kernel_id, self._A1.dtype, self._A1.shape[1], _n_cols(x))
        if layout == ops.LAYOUT_TC and not self._cache.tc_ok(self._kernel_id):
            return ops.LAYOUT_SIMT
        return layout

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        P1, P2 = self._cache.get(self._layout_for(x))
        return ops.matmat_packed(P1, P2, x, self._kernel_id, self._scaling)

    def _adjoint(self, x: torch.Tensor) -> torch.Tensor:
        # K(A1, A2)^T = K(A2, A1): same kernel with the operands' roles swapped
        P1, P2 = self._cache.get(self._layout_for(x))
        return ops.matmat_packed(P2, P1, x, self._kernel_id, self._scaling)

    # -- fused products (rlaopt_b200.linops.apply_fused) -----------------------------
    @staticmethod
    def fused_reductions_ok(k: int, gram_cols: int = 0) -> bool:
        return ops.fused_reductions_supported(k, gram_cols)

    def matmat_fused(self, x: torch.Tensor, **epilogue):
        """``alpha c K x + beta addend[..] + gamma rhs[..]`` with optional Gram / column norms, one pass
        (``ops.matmat_packed_fused``): ``A P + reg P`` with ``P^T A P``, residuals with their norms."""
        P1, P2 = self._cache.get(self._layout_for(x))
        return ops.matmat_packed_fused(P1, P2, x, self._kernel_id, self._scaling, **epilogue)

    # -- oracles --------------------------------------------------------------
    def _oracle_packs(self, kind: str, blk: torch.Tensor, x: torch.Tensor):
        """Packs of ``A1[blk]`` (and ``A2[blk]``), memoised on the identity of ``blk``.

        SAP asks for ``A_blk_oracle(blk)`` once per power-iteration matvec with the
        same ``blk`` object (``rlaopt/solvers/sap.py:96-97``); the memo turns those
        repeats into cache hits.  The blocks are packed with the operator's center, so the
        operator's norm bound covers them (no read-back per block when ``A1 is A2``).
        """
        cache = self._cache
        cache.validate()
        layout = ops.choose_layout(self._kernel_id, self._A1.dtype, self._A1.shape[1], _n_cols(x))
        key = f"{kind}:{layout}"
        memo = self._oracle_memo.get(key)
        if memo is not None and memo[0] is blk and memo[1] == (blk._version, cache._state):
            return memo[2]
        ls = self._kernel_config.lengthscale
        packs = None
        if layout == ops.LAYOUT_TC:
            c = cache.center()
            Pr = ops.pack_points(self._A1, ls, blk, layout, c)
            if kind == "row":
                Pc = cache.cols(layout)
            else:
                Pc = Pr if cache.shared else ops.pack_points(self._A2, ls, blk, layout, c)
            rmax, cmax = cache.known_sqnorm("rows"), cache.known_sqnorm("cols")
            if rmax is None or cmax is None or not ops.tc_accuracy_ok(self._kernel_id, rmax, cmax):
                # no operator-wide bound (or it fails): this block's own norms decide
                rmax, cmax = Pr.max_sqnorm, Pc.max_sqnorm
            if ops.tc_accuracy_ok(self._kernel_id, rmax, cmax):
                packs = (Pr, Pc)
            else:
                layout = ops.LAYOUT_SIMT
        if packs is None:
            Pr = ops.pack_points(self._A1, ls, blk, layout)
            if kind == "row":
                packs = (Pr, cache.cols(layout))
            else:
                packs = (Pr, Pr if cache.shared else ops.pack_points(self._A2, ls, blk, layout))
        self._oracle_memo[key] = (blk, (blk._version, cache._state), packs)
        return packs

    def _get_kernel_linop(self, kind: str, blk: torch.Tensor) -> LinOp:
        if not isinstance(blk, torch.Tensor) or blk.ndim != 1:
            raise ValueError("blk must be a 1D index tensor")
        n_cols = blk.shape[0] if kind == "blk" else self._A2.shape[0]

        def matvec(x: torch.Tensor) -> torch.Tensor:
            Pr, Pc = self._oracle_packs(kind, blk, x)
            return ops.matmat_packed(Pr, Pc, x, self._kernel_id, self._scaling)

        def matmat_fused(x: torch.Tensor, **epilogue):
            Pr, Pc = self._oracle_packs(kind, blk, x)
            return ops.matmat_packed_fused(Pr, Pc, x, self._kernel_id, self._scaling, **epilogue)

        op = LinOp(
            device=self.device,
            shape=torch.Size((blk.shape[0], n_cols)),
            matvec=matvec,
            matmat=matvec,
            dtype=self.dtype,
        )
        # the block gradient of SAP / ASkotch in one pass: K[blk, :] Y + reg Y[blk] - B[blk] (sap.py:113-127)
        op.matmat_fused = matmat_fused
        op.fused_reductions_ok = ops.fused_reductions_supported
        return op

    def row_oracle(self, blk: torch.Tensor) -> LinOp:
        """``c * K(A1[blk], A2)`` (``rlaopt/kernels/base.py:124-125``); forward products only."""
        return self._get_kernel_linop("row", blk)

    def blk_oracle(self, blk: torch.Tensor) -> LinOp:
        """``c * K(A1[blk], A2[blk])`` (``rlaopt/kernels/base.py:127-128``)."""
        return self._get_kernel_linop("blk", blk)

    def _clear_cache(self) -> None:
        self._cache.clear()
        self._oracle_memo.clear()


def _device_sort_key(dev: torch.device):
    return (dev.type, -1 if dev.index is None else dev.index)


class _DistributedKernelLinOp(DistributedTwoSidedLinOp, ScaleMixin):
    """Kernel operator with the rows of ``A1`` partitioned over ``devices``.

    Partitioning follows the reference exactly: ``torch.chunk(arange(n), g)`` for
    the operator and for the row oracle's column split, ``torch.chunk(arange(b), g)``
    for the block oracle (``rlaopt/kernels/base.py:297-302,462``).  ``A2`` is
    replicated on every device when ``use_full_kernel`` (``base.py:143-144``); with
    ``use_full_kernel=False`` only the ``A2`` chunks are resident and the operator
    serves oracles only (``base.py:311-316,383-406``).
    """

    def __init__(
        self,
        A1: torch.Tensor,
        A2: torch.Tensor,
        kernel_config: KernelConfig,
        devices: set[torch.device],
        use_full_kernel: bool,
        _kernel_key: str,
    ):
        self._check_inputs(A1, A2, kernel_config, devices)
        self._A1, self._A2 = A1, A2
        self._kernel_config = kernel_config
        self._kernel_key = _kernel_key
        self._kernel_id = ops.kernel_id(_kernel_key)
        self._initialize_scaling(getattr(kernel_config, "const_scaling", 1.0))

        # a set has no order; sort for a deterministic device <-> chunk assignment
        ordered = sorted(devices, key=_device_sort_key)
        self._ordered_devices = ordered
        self._kernel_config_devices = {dev: kernel_config.to(dev) for dev in ordered}

        g = len(ordered)
        self.A1_row_chunks = torch.chunk(torch.arange(A1.shape[0]), g, dim=0)
        self.A2_row_chunks = torch.chunk(torch.arange(A2.shape[0]), g, dim=0)

        # A2 column blocks, one per device (used by the row oracle)
        self.A2_chunks = [
            A2[chunk[0] : chunk[-1] + 1].to(dev) for dev, chunk in zip(ordered, self.A2_row_chunks)
        ]
        self._A2_chunk_packs: list[dict] = [dict() for _ in self.A2_chunks]

        shared = _same_storage(A1, A2)
        kernel_ops = []
        for dev, chunk in zip(ordered, self.A1_row_chunks):
            lo, hi = int(chunk[0]), int(chunk[-1]) + 1
            if use_full_kernel:
                A2_dev = A2.to(dev)
                A1_dev = A2_dev[lo:hi] if shared else A1[lo:hi].to(dev)
                kernel_ops.append(
                    _KernelLinOp(A1_dev, A2_dev, self._kernel_config_devices[dev], _kernel_key=_kernel_key)
                )
            else:
                # shape-only placeholders: the operator then serves oracles only
                kernel_ops.append(
                    TwoSidedLinOp(
                        device=dev,
                        shape=torch.Size((hi - lo, A2.shape[0])),
                        matvec=_not_materialised,
                        rmatvec=_not_materialised,
                        matmat=_not_materialised,
                        rmatmat=_not_materialised,
                        dtype=A1.dtype,
                    )
                )

        super().__init__(
            shape=torch.Size((A1.shape[0], A2.shape[0])),
            A=kernel_ops,
            distribution_mode="row",
        )
        self.kernel_ops = kernel_ops

    @property
    def A1(self) -> torch.Tensor:
        return self._A1

    @property
    def A2(self) -> torch.Tensor:
        return self._A2

    @property
    def kernel_config(self) -> KernelConfig:
        return self._kernel_config

    def _check_inputs(self, A1: Any, A2: Any, kernel_config: Any, devices: Any) -> None:
        _is_torch_tensor(A1, "A1")
        _is_torch_tensor(A2, "A2")
        if A1.ndim != 2:
            raise ValueError(f"A must be a 2D tensor, got {A1.ndim}D tensor.")
        if A2.ndim != 2:
            raise ValueError(f"A must be a 2D tensor, got {A2.ndim}D tensor.")
        if A1.dtype != A2.dtype:
            raise ValueError("A1 and A2 must have the same dtype.")
        if A1.shape[1] != A2.shape[1]:
            raise ValueError(
                f"A1 and A2 must have the same number of features, got {A1.shape[1]} and {A2.shape[1]}."
            )
        _is_kernel_config(kernel_config, "kernel_config")
        _is_set(devices, "devices")
        if len(devices) == 0:
            raise ValueError("devices must be a non-empty set.")
        if not all(isinstance(d, torch.device) for d in devices):
            raise ValueError("All elements in devices must be torch.device instances.")

    # -- oracles ------------------------------------------------------------
    def _chunk_center(self, i: int) -> torch.Tensor:
        """Shift of the tensor-core packs on device ``i``: the column means of its resident A2 chunk.  Any vector
        works as long as the row and column operand of one product share it; partial products are summed after."""
        cache = self._A2_chunk_packs[i]
        if "center" not in cache:
            cache["center"] = ops.column_mean(self.A2_chunks[i])
        return cache["center"]

    def _chunk_pack(self, i: int, layout: int):
        """Pack of the i-th A2 column block on its device (cached across oracle calls)."""
        cache = self._A2_chunk_packs[i]
        if layout not in cache:
            dev = self.A2_chunks[i].device
            cache[layout] = ops.pack_points(
                self.A2_chunks[i], self._kernel_config_devices[dev].lengthscale, None, layout,
                self._chunk_center(i) if layout == ops.LAYOUT_TC else None,
            )
        return cache[layout]

    def _gather_rows(self, A: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        return A[idx.to(A.device)]

    def row_oracle(self, blk: torch.Tensor) -> _DistributedLinOp:
        """``c * K(A1[blk], A2)``, column-partitioned over the A2 chunks (partials summed)."""
        A1b = self._gather_rows(self._A1, blk)
        d, dtype, kid, scale = self._A1.shape[1], self._A1.dtype, self._kernel_id, self._scaling
        row_ops = []
        for i, A2_chunk in enumerate(self.A2_chunks):
            dev = A2_chunk.device
            fn = _ShardProduct(self, i, dev, A1b, None, kid, scale, d, dtype)
            row_ops.append(
                LinOp(
                    device=dev,
                    shape=torch.Size((blk.shape[0], A2_chunk.shape[0])),
                    matvec=fn,
                    matmat=fn,
                    dtype=dtype,
                )
            )
        return _DistributedLinOp(
            shape=torch.Size((blk.shape[0], self._A2.shape[0])),
            A=row_ops,
            distribution_mode="column",
            is_new=False,
        )

    def blk_oracle(self, blk: torch.Tensor) -> _DistributedLinOp:
        """``c * K(A1[blk], A2[blk])``, row-partitioned over chunks of ``blk`` (concatenated)."""
        A1b = self._gather_rows(self._A1, blk)
        A2b = A1b if _same_storage(self._A1, self._A2) else self._gather_rows(self._A2, blk)
        d, dtype, kid, scale = self._A1.shape[1], self._A1.dtype, self._kernel_id, self._scaling
        blk_chunks = torch.chunk(torch.arange(blk.shape[0]), len(self._ordered_devices), dim=0)
        block_ops = []
        for i, (dev, pos) in enumerate(zip(self._ordered_devices, blk_chunks)):
            lo, hi = int(pos[0]), int(pos[-1]) + 1
            fn = _ShardProduct(self, i, dev, A1b[lo:hi], A2b, kid, scale, d, dtype)
            block_ops.append(
                LinOp(device=dev, shape=torch.Size((hi - lo, blk.shape[0])), matvec=fn, matmat=fn, dtype=dtype)
            )
        return _DistributedLinOp(
            shape=torch.Size((blk.shape[0], blk.shape[0])),
            A=block_ops,
            distribution_mode="row",
            is_new=False,
        )

    def shutdown(self) -> None:
        """Drop the cached packs and close the operator (``rlaopt/kernels/base.py:507-520``)."""
        for op in getattr(self, "kernel_ops", []):
            if hasattr(op, "_clear_cache"):
                op._clear_cache()
        for cache in getattr(self, "_A2_chunk_packs", []):
            cache.clear()
        super().shutdown()


def _not_materialised(x):
    raise RuntimeError(
        "this distributed kernel operator was built with use_full_kernel=False: "
        "only row_oracle / blk_oracle are available"
    )


class _ShardProduct:
    """One device's share of an oracle product: ``c * K(rows, cols) @ x`` on device ``i``.

    ``rows`` / ``cols`` are point sets living anywhere; they are moved to the shard's
    device and packed on first use.  ``cols=None`` means "the i-th resident A2 chunk".
    """

    def __init__(self, owner, i, dev, rows, cols, kid, scale, d, dtype):
        self.owner, self.i, self.dev, self.rows, self.cols = owner, i, dev, rows, cols
        self.kid, self.scale, self.d, self.dtype = kid, scale, d, dtype
        self._packs: dict[int, tuple] = {}

    def _build(self, layout: int):
        owner, i, dev = self.owner, self.i, self.dev
        ls = owner._kernel_config_devices[dev].lengthscale
        rows = self.rows.to(dev)
        tc = layout == ops.LAYOUT_TC
        if self.cols is None:
            Pc = owner._chunk_pack(i, layout)
            Pr = ops.pack_points(rows, ls, None, layout, Pc.center)
        elif self.cols is self.rows:
            Pr = Pc = ops.pack_points(rows, ls, None, layout, ops.column_mean(rows) if tc else None)
        else:
            cols = self.cols.to(dev)
            c = ops.column_mean(cols) if tc else None
            Pr = ops.pack_points(rows, ls, None, layout, c)
            Pc = ops.pack_points(cols, ls, None, layout, c)
        return Pr, Pc

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        layout = ops.choose_layout(self.kid, self.dtype, self.d, _n_cols(x))
        if layout not in self._packs:
            Pr, Pc = self._build(layout)
            if layout == ops.LAYOUT_TC and not ops.tc_accuracy_ok(self.kid, Pr.max_sqnorm, Pc.max_sqnorm):
                Pr, Pc = self._build(ops.LAYOUT_SIMT)  # norms beyond the tensor-core accuracy budget
            self._packs[layout] = (Pr, Pc)
        Pr, Pc = self._packs[layout]
        return ops.matmat_packed(Pr, Pc, x, self.kid, self.scale)
