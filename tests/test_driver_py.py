"""acg_b200.driver: the multi-process command line (read, partition, solve, report)
that stands in for the MPI part of cuda/acg-cuda.c.  CPU: --dry-run under gloo
(ingest + decomposition for every partition source); GPU: real solves."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from acg_b200 import dist as abdist
from acg_b200 import matgen as mg
from acg_b200 import mtxio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, argv, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), "-m", "acg_b200.driver"] + argv
    env = dict(os.environ, OMP_NUM_THREADS="2", PYTHONPATH=ROOT)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def _decomposition(stderr):
    rows = {}
    for m in re.finditer(r"^\s+(\d+): (\d+) (\d+) (\d+) (\d+) (\d+) (\d+)$", stderr, re.M):
        rows[int(m.group(1))] = tuple(int(g) for g in m.groups()[1:])
    return rows


def test_rowparts_file_roundtrip(tmp_path):
    rp = np.array([0, 2, 1, 1, 0, 2], np.int32)
    path = str(tmp_path / "parts.mtx")
    mtxio.write_rowparts(path, rp)
    assert open(path).read().split("\n")[:3] == ["%%MatrixMarket vector array integer general", "6", "1"]
    assert np.array_equal(mtxio.read_rowparts(path), rp)


@pytest.mark.parametrize("partition", ["rows", "nnz", "metis", "file"])
def test_dry_run_decomposition(partition, ab, tmp_path):
    """Three processes read their parts from the binary file; the printed decomposition
    equals what partitioning the whole matrix gives, and the communication matrix file
    is the one assembled from the parts."""
    n, r, c, v = mg.stencil3d_27pt(9, 8, 7)
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=True)
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v)
    if partition == "rows":
        rowparts, arg = abdist.contiguous_partition(n, 3), "rows"
    elif partition == "nnz":
        rowparts, arg = abdist.balanced_rows_partition(A, 3), "nnz"
    elif partition == "metis":
        rowparts, arg = A.partition_rows(3, kway=False, seed=1)[0], "metis"
    else:
        rowparts = (np.arange(n) * 7 % 3).astype(np.int32)
        arg = str(tmp_path / "parts.mtx")
        mtxio.write_rowparts(arg, rowparts)
    cm = str(tmp_path / "comm.mtx")
    p = _run(3, [path, "--binary", "--partition", arg, "--dry-run", "--output-comm-matrix", cm])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    got = _decomposition(p.stderr)
    parts = A.partition(3, rowparts)
    M = np.stack([m.comm_matrix_row(3) for m in parts])
    for q, m in enumerate(parts):
        m.dsymv_init(0.0)
        assert got[q] == (m.c.nownedrows, m.c.ninnerrows, m.c.nborderrows, m.c.nghostrows,
                          m.c.fnpnzs + m.c.onpnzs, int(M[q].sum()))
    want = str(tmp_path / "want.mtx")
    mtxio.write_comm_matrix(want, M)
    assert open(cm).read() == open(want).read()


def test_refuses_to_solve_without_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    n, r, c, v = mg.laplace3d_7pt(5)
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=False)
    p = subprocess.run([sys.executable, "-m", "acg_b200.driver", path], capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT, timeout=120)
    assert p.returncode != 0 and "no CUDA device" in p.stderr


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,solver,method", [(1, "acg", "cg"), (1, "acg-pipelined", "cg_pipelined"),
                                                 (2, "acg-pipelined", "cg_pipelined"), (2, "acg-device", "cg")])
def test_solves_from_file(nproc, solver, method, oracle, tmp_path):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    n, r, c, v = mg.stencil3d_27pt(16)
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=True)
    sol = str(tmp_path / "x.mtx")
    p = _run(nproc, [path, "--binary", "--solver", solver, "--max-iterations", "300", "--residual-rtol", "1e-9",
                     "--warmup", "2", "--output-solution", sol, "--manufactured-solution", "--seed", "5"])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    its = int(re.search(r"^\s*iterations: (\d+)", p.stderr, re.M).group(1))
    e0 = float(re.search(r"^initial error 2-norm: (\S+)", p.stderr, re.M).group(1))
    e1 = float(re.search(r"^error 2-norm: (\S+)", p.stderr, re.M).group(1))
    assert e0 == pytest.approx(1.0, rel=1e-12) and e1 < 1e-7
    # same system on the oracle: b = A x* with the driver's x*
    xs = np.random.default_rng(5).uniform(-1.0, 1.0, n); xs /= np.linalg.norm(xs)
    csr = oracle.full_csr(n, r, c, v)
    b = oracle.dsymv(csr, 1.0, xs, 0.0, np.zeros(n))
    want = getattr(oracle, method)(csr, b, maxits=300, rtol=1e-9)
    assert abs(its - want["niterations"]) <= 1            # b differs by rounding (scipy vs oracle summation order)
    x = np.array([float(t) for t in open(sol).read().split("\n")[2:] if t])
    assert np.abs(x - want["x"]).max() <= 1e-8 * np.abs(want["x"]).max()


def test_manufactured_rhs_of_a_part_equals_global_product(ab, oracle):
    """driver.local_rhs: b = A x* computed part by part (local block + border x ghost block on
    the owned and ghost entries of x*) equals the rows of the global product."""
    import types
    from acg_b200 import driver
    n, r, c, v = mg.rmat_spd(3000, 20000, seed=2)
    csr = oracle.full_csr(n, r, c, v)
    xs = np.random.default_rng(4).uniform(-1.0, 1.0, n); xs /= np.linalg.norm(xs)
    want = oracle.dsymv(csr, 1.0, xs, 0.0, np.zeros(n))
    args = types.SimpleNamespace(manufactured_solution=True, seed=4)
    parts = ab.SymCsrMatrix.init_real_double(n, r, c, v).partition(4, (np.arange(n) * 13 % 4).astype(np.int32))
    for m in parts:
        m.dsymv_init(0.0)
        b, xloc = driver.local_rhs(args, m, n)
        own = m.nzrows[:m.c.nownedrows]
        assert np.array_equal(xloc, xs[own])
        assert np.abs(b - want[own]).max() <= 1e-13 * np.abs(want).max()
    whole = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    b, _ = driver.local_rhs(args, whole, n)
    assert np.abs(b - want).max() <= 1e-13 * np.abs(want).max()
    b1, none = driver.local_rhs(types.SimpleNamespace(manufactured_solution=False, seed=0), whole, n)
    assert none is None and np.all(b1 == 1.0)
