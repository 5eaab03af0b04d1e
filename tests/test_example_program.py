"""examples/acgb200_solve.c: a C program linked against libacgb200.so alone (no
reference code, no Python): Matrix Market in, solver report out.  On a machine
without a GPU it must fail loudly at acgsolvercuda_init (there is no CPU path)."""
import os
import re
import subprocess

import numpy as np
import pytest

from acg_b200 import matgen as mg
from acg_b200 import mtxio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "acgb200_solve")


def _build():
    p = subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "warning" not in p.stderr


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    _build()
    assert subprocess.run([EXE, "--help"], capture_output=True, text=True).stdout.startswith("usage: acgb200_solve")
    if _has_gpu():
        pytest.skip("GPU present: the run is covered by the gpu test")
    n, r, c, v = mg.stencil3d_27pt(5)
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=True)
    p = subprocess.run([EXE, path, "--binary"], capture_output=True, text=True)
    assert p.returncode == 1 and "acgsolvercuda_init" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("solver,method,binary", [("acg", "cg", True), ("acg-pipelined", "cg_pipelined", False)])
def test_example_solves(solver, method, binary, oracle, tmp_path):
    if not os.path.exists(EXE):
        _build()
    n, r, c, v = mg.stencil3d_27pt(14)
    csr = oracle.full_csr(n, r, c, v)
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=binary)
    p = subprocess.run([EXE, path] + (["--binary"] if binary else []) +
                       ["--solver", solver, "--max-iterations", "200", "--residual-rtol", "1e-9", "--warmup", "1",
                        "--print-solution"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    its = int(re.search(r"^\s*iterations: (\d+)", p.stderr, re.M).group(1))
    want = getattr(oracle, method)(csr, np.ones(n), maxits=200, rtol=1e-9)
    assert its == want["niterations"]
    x = np.array([float(t) for t in p.stdout.splitlines()[2:]])
    assert len(x) == n and np.abs(x - want["x"]).max() <= 1e-9 * np.abs(want["x"]).max()
