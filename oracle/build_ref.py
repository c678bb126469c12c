"""Install the UNMODIFIED reference into ``baseline/_ref`` (git-ignored; it travels to the GPU box with the snapshot).

    python oracle/build_ref.py            # no-op when baseline/_ref/rlaopt already exists

Recipe of the bench contract: ``pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
--target baseline/_ref <copy of /root/reference>`` with ``RLAOPT_CPU_ONLY=1 RLAOPT_USE_OPENMP=0`` (the reference's
own build switches, ``setup.py:23-36``; the first ``g++`` on PATH here cannot link ``-fopenmp``) from a scratch copy,
because the build writes into the source tree and ``/root/reference`` is read-only.  ``--no-deps``: PyKeOps / wandb
cannot be resolved offline.  The result imports everything except ``rlaopt.kernels`` (``import pykeops`` fails), i.e.
the reference's own ``LinSys`` / ``PCG`` / ``SAP`` / ``Nystrom`` / sketches run, over an operator the caller supplies.

Test / bench infrastructure only: ``bench.py --impl reference`` and ``oracle/gen_solver_golden.py`` import it; the
product never does.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
REFERENCE = os.environ.get("RLAOPT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(TARGET, "rlaopt", "__init__.py"))


def build(force: bool = False) -> bool:
    """Returns True when baseline/_ref holds an importable reference afterwards."""
    if available() and not force:
        return True
    if not os.path.isdir(REFERENCE):
        return False
    tmp = tempfile.mkdtemp(prefix="rlaopt_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git"))
        if force and os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET, exist_ok=True)
        env = dict(os.environ, RLAOPT_CPU_ONLY="1", RLAOPT_USE_OPENMP="0")
        if os.path.exists("/usr/bin/g++"):
            env.update(CXX="/usr/bin/g++", CC="/usr/bin/gcc")
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        proc = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout[-2000:] + proc.stderr[-2000:])
            return False
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print(TARGET if ok else "reference not installed")
    sys.exit(0 if ok else 1)
