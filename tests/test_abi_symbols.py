"""The C-ABI library loads and exports every symbol ``include/rlaopt_b200.h`` declares.

No compute call is made (there is no GPU in the CPU tier); only pure-host entry
points are invoked.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rlaopt_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rlaopt_b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from rlaopt_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from rlaopt_b200.csrc import build

        build.build()
    return _lib.load()


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for required in (
        "rlaopt_b200_abi_version",
        "rlaopt_b200_pack_points_f32",
        "rlaopt_b200_pack_points_f64",
        "rlaopt_b200_matmat_packed_f32",
        "rlaopt_b200_matmat_packed_f64",
        "rlaopt_b200_kernel_matmat_f32",
        "rlaopt_b200_kernel_matmat_f64",
        "rlaopt_b200_kernel_matmat_host_f32",
        "rlaopt_b200_column_mean_f32",
        "rlaopt_b200_packed_stats_host",
    ):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(lib._name)
    for name in _declared_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_python_prototypes_cover_the_header(lib):
    from rlaopt_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == _declared_functions()


def test_host_only_entry_points(lib):
    from rlaopt_b200._lib import LAYOUT_SIMT, LAYOUT_TC

    assert lib.rlaopt_b200_abi_version() == 2
    # SIMT layout: rows padded to 128, features to 8
    assert lib.rlaopt_b200_packed_bytes(10, 3, 4, LAYOUT_SIMT) == 128 * 8 * 4
    assert lib.rlaopt_b200_packed_bytes(129, 9, 8, LAYOUT_SIMT) == 256 * 16 * 8
    assert lib.rlaopt_b200_packed_bytes(0, 3, 4, LAYOUT_SIMT) == 0
    # every kernel / dtype is supported on the CUDA-core layout
    for kid in range(5):
        for elem in (4, 8):
            assert lib.rlaopt_b200_layout_supported(kid, elem, 3, 1, LAYOUT_SIMT) == 1
    assert lib.rlaopt_b200_layout_supported(7, 4, 3, 1, LAYOUT_SIMT) == 0
    # Laplace (L1) and fp64 never run on the tensor-core layout
    assert lib.rlaopt_b200_layout_supported(1, 4, 128, 64, LAYOUT_TC) == 0
    assert lib.rlaopt_b200_layout_supported(0, 8, 128, 64, LAYOUT_TC) == 0
    # the four L2-distance kernels in fp32: tensor-core layout up to d = 2048 (X in TMEM up to 192, K-block streaming beyond)
    for kid in (0, 2, 3, 4):
        for d in (1, 128, 192, 193, 784, 2048):
            assert lib.rlaopt_b200_layout_supported(kid, 4, d, 1, LAYOUT_TC) == 1
        assert lib.rlaopt_b200_layout_supported(kid, 4, 2049, 1, LAYOUT_TC) == 0
    # TC pack: 256 B header + one fp32 norm per padded point + (hi | lo) fp16 images of 64 points x 64-feature K-blocks
    assert lib.rlaopt_b200_packed_bytes(10, 3, 4, LAYOUT_TC) == 256 + 128 * 4 + 2 * (2 * 64 * 128)
    assert lib.rlaopt_b200_packed_bytes(129, 130, 4, LAYOUT_TC) == 256 + 256 * 4 + 4 * (2 * 3 * 64 * 128)
    assert lib.rlaopt_b200_packed_bytes(10, 3, 8, LAYOUT_TC) == 0


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    from rlaopt_b200 import _lib

    rc = lib.rlaopt_b200_matmat_packed_f32(None, 4, None, 4, 3, None, 1, 1, None, 1, 99, 1.0, 0, None, 0, None)
    assert rc == -1
    assert b"unknown kernel" in lib.rlaopt_b200_last_error()
    rc = lib.rlaopt_b200_pack_points_f32(None, 4, 4, 0, 0, None, 1.0, None, None, 0, None, None)
    assert rc == -1
    # identity gather of more points than the source holds
    rc = lib.rlaopt_b200_pack_points_f32(None, 5, 4, 3, 3, None, 1.0, None, None, 0, None, None)
    assert rc == -1 and b"n_src" in lib.rlaopt_b200_last_error()
    # statistics exist for tensor-core packs only; the column-mean workspace is one fp64 row per 4096 points
    assert lib.rlaopt_b200_packed_stats_host(None, 0, None, None, None) == -3
    assert lib.rlaopt_b200_column_mean_workspace_bytes(4097, 5) == 2 * 5 * 8
    assert lib.rlaopt_b200_column_mean_f32(None, 4, 4, 3, 3, None, None, None, 0, None) == -1
    with pytest.raises(RuntimeError, match="code -1"):
        _lib.check(rc, "pack_points")


def test_register_contraction_workspace_plan(lib, monkeypatch):
    """k <= 4 on the tensor-core layout: the V workspace is one {V tile, column norms} record of (kv + 1) x 256 B per
    64-column sub-tile (kv = 1, 2, 4) instead of the fp16 hi/lo record of 16 x 256 + 16 + 256 B; RLAOPT_B200_TC_KV=0 switches
    back.  Host-only: the plan does not touch the GPU."""
    from rlaopt_b200._lib import LAYOUT_TC

    n, m, d = 1000, 64 * 500 + 1, 16  # one wave: no column splits, no partial buffers
    sub_tiles = 501
    monkeypatch.delenv("RLAOPT_B200_TC_KV", raising=False)
    for k, kv in ((1, 1), (2, 2), (3, 4), (4, 4)):
        assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, d, k, 4, LAYOUT_TC) == sub_tiles * (kv + 1) * 256
    image = -(-sub_tiles * (16 * 256 + 16 + 64 * 4) // 256) * 256  # fp16 hi | lo image, 1 / s trailer, 64 column norms
    assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, d, 5, 4, LAYOUT_TC) == image
    monkeypatch.setenv("RLAOPT_B200_TC_KV", "0")
    assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, d, 1, 4, LAYOUT_TC) == image
    # wide d (X streamed through smem) keeps the MMA2 path
    monkeypatch.delenv("RLAOPT_B200_TC_KV", raising=False)
    assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, 500, 1, 4, LAYOUT_TC) >= image  # + split-column partials
    # 128 < d <= 192: three K-blocks leave ring depths that are ambiguous for three warpgroups -- MMA2 path (tc_plan)
    assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, 150, 1, 4, LAYOUT_TC) >= image
    assert lib.rlaopt_b200_matmat_workspace_bytes(n, m, 128, 1, 4, LAYOUT_TC) == sub_tiles * 2 * 256


def test_two_chunk_knob_does_not_change_the_workspace_plan(lib, monkeypatch):
    """k > 128: whether one CTA contracts one or two 128-column chunks of V per P' (RLAOPT_B200_TC_DUAL) is decided at launch;
    the V records and the split-column partials are laid out the same way, so a workspace sized under one setting serves
    every other (the knob is read per launch).  Host-only."""
    from rlaopt_b200._lib import LAYOUT_TC

    shapes = [(37888, 1_000_000, 128, 256), (94720, 2_000_000, 64, 1000), (1000, 777, 64, 130), (513, 129, 100, 300)]
    for n, m, d, k in shapes:
        sizes = set()
        for dual in (None, "0", "1", "2", "3"):
            if dual is None:
                monkeypatch.delenv("RLAOPT_B200_TC_DUAL", raising=False)
            else:
                monkeypatch.setenv("RLAOPT_B200_TC_DUAL", dual)
            sizes.add(lib.rlaopt_b200_matmat_workspace_bytes(n, m, d, k, 4, LAYOUT_TC))
        assert len(sizes) == 1 and sizes.pop() > 0, (n, m, d, k)


def test_torch_library_op_is_registered_from_cpp(lib):
    """``torch.ops.rlaopt.kernel_matmat``: schema defined by TORCH_LIBRARY_FRAGMENT(rlaopt, m) in csrc/torch_op.cpp (the
    reference's registration pattern, rlaopt/csrc/cpp/csc_matmat.cpp:83-87); the CPU key raises -- no CPU fallback."""
    import torch

    from rlaopt_b200 import ops
    from rlaopt_b200.csrc import build

    build.build_torch_op()
    assert ops.load_torch_op()
    schema = str(torch.ops.rlaopt.kernel_matmat.default._schema)
    assert schema.startswith("rlaopt::kernel_matmat(Tensor A1, Tensor A2, Tensor V, int kernel_id, float lengthscale, "
                             "Tensor? lengthscale_vec, float const_scaling, bool transpose=False, Tensor? row_idx=None, "
                             "Tensor? col_idx=None) -> Tensor")
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        torch.ops.rlaopt.kernel_matmat(torch.randn(4, 3), torch.randn(5, 3), torch.randn(5, 2), 0, 1.0, None, 1.0)
    # the Python-registered twin of round 1 raises the same way
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        torch.ops.rlaopt_b200.kernel_matmat(torch.randn(4, 3), torch.randn(5, 3), torch.randn(5, 2), 0, 1.0, None, 1.0)


def test_header_is_plain_c_and_links(lib, tmp_path):
    """The boundary is a C ABI: ``include/rlaopt_b200.h`` compiles as pedantic C99 (no C++ / torch / CUDA types in the
    signatures) and a C program links against the shared library and calls a host-only entry point."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    if not os.path.exists(gcc):
        pytest.skip("no C compiler")
    from rlaopt_b200 import _lib

    src = tmp_path / "abi.c"
    src.write_text(
        '#include "rlaopt_b200.h"\n'
        "int main(void) {\n"
        "    rlaopt_b200_epilogue_f32 e = {0};\n"
        "    (void)e;\n"
        "    if (rlaopt_b200_layout_supported(RLAOPT_B200_KERNEL_LAPLACE, 4, 32, 16, RLAOPT_B200_LAYOUT_TC)) return 2;\n"
        "    if (rlaopt_b200_packed_bytes(10, 3, 4, RLAOPT_B200_LAYOUT_SIMT) != 128 * 8 * 4) return 3;\n"
        "    return rlaopt_b200_abi_version() == RLAOPT_B200_ABI_VERSION ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    proc = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{os.path.join(ROOT, 'include')}", str(src),
                           "-o", str(exe), f"-L{libdir}", "-lrlaopt_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    assert subprocess.run([str(exe)]).returncode == 0
