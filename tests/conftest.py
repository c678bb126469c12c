"""Shared pytest configuration.

* ``gpu`` marker: tests that need a CUDA device (run on the B200 box with ``-m gpu``).
* ``oracle_backend`` fixture: swaps the two CUDA entry points of ``rlaopt_b200.ops``
  for stand-ins built on the CPU oracle, so the *host-side* logic (shape dispatch,
  oracles, multi-device partitioning) can be exercised without a GPU.  The oracle
  is test infrastructure; the product never imports it.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "kernels_ref.pt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_cases():
    return torch.load(GOLDEN, weights_only=False)["cases"]


class _FakePack:
    def __init__(self, X, lengthscale, layout):
        self.X, self.lengthscale, self.layout = X, lengthscale, layout
        self.n, self.d, self.dtype, self.device = X.shape[0], X.shape[1], X.dtype, X.device
        self.buf = X
        self.center = None
        self.max_sqnorm = 0.0


@pytest.fixture
def oracle_backend(monkeypatch):
    """Route ops.pack_points / ops.matmat_packed through the CPU oracle (tests only)."""
    from oracle import kernel_oracle as ko
    from rlaopt_b200 import ops

    calls = {"pack": 0, "matmat": 0}

    def fake_pack(X, lengthscale, idx=None, layout=0, center=None):
        calls["pack"] += 1
        Xg = X if idx is None else X[idx.to(X.device)]
        return _FakePack(Xg, lengthscale, layout)

    def fake_matmat(rows, cols, V, kernel, const_scaling=1.0):
        calls["matmat"] += 1
        assert rows.device == cols.device == V.device, "operands must be co-located"
        if V.shape[0] != cols.n:
            raise ValueError("dimension mismatch")
        return ko.kernel_matmat(rows.X, cols.X, V, ops.kernel_id(kernel), rows.lengthscale, const_scaling)

    monkeypatch.setattr(ops, "pack_points", fake_pack)
    monkeypatch.setattr(ops, "matmat_packed", fake_matmat)
    monkeypatch.setattr(ops, "choose_layout", lambda *a, **k: 0)
    return calls
