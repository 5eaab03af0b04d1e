"""Full-storage expansion on the device (acg_b200/csrc/expand.cu) against the
host routine acgsymcsrmatrix_dsymv_init, which tests/test_host_structs.py pins byte for byte to the
reference build: every array must be identical -- row pointers, column order inside each row,
values, and the border x ghost block -- for whole matrices and for the parts of partitions (local
renumbering puts transposed entries on both sides of a row's own entries), with a diagonal shift and
with 1-based indices.  And a solver built on a matrix WITHOUT full storage (acgsolvercuda_init
expands on the device itself) must give the oracle's iterates."""
import numpy as np
import pytest

from acg_b200 import matgen as mg
from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

ARRAYS = ("frowptr", "fcolidx", "fa", "orowptr", "ocolidx", "oa")


def _same(host, dev):
    assert host.c.fnpnzs == dev.c.fnpnzs and host.c.onpnzs == dev.c.onpnzs
    for k in ARRAYS:
        a, b = getattr(host, k), getattr(dev, k)
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert a.tobytes() == b.tobytes(), k


CASES = [("27pt-12", lambda: mg.stencil3d_27pt(12)), ("7pt-aniso", lambda: mg.laplace3d_7pt(31, 17, 23)),
         ("1d5pt", lambda: mg.poisson1d_5pt(20000)), ("rand", lambda: mg.random_spd(400, 0.5, 2)),
         ("rmat-hubs", lambda: mg.rmat_spd(30000, 600000, seed=8)), ("n1", lambda: mg.poisson1d_3pt(1)),
         ("n3", lambda: mg.poisson1d_3pt(3))]


@pytest.mark.parametrize("eps", [0.0, 0.375])
@pytest.mark.parametrize("name,gen", CASES, ids=[c[0] for c in CASES])
def test_whole_matrix_byte_identical(name, gen, eps, ab):
    n, r, c, v = gen()
    host = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(eps)
    dev = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init_cuda(eps)
    _same(host, dev)


@pytest.mark.parametrize("name,gen,nparts,kind", [("27pt-16-block8", lambda: mg.stencil3d_27pt(16), 8, "block"),
                                                  ("7pt-20-contig5", lambda: mg.laplace3d_7pt(20), 5, "contiguous"),
                                                  ("rmat-random4", lambda: mg.rmat_spd(20000, 300000, seed=3), 4, "random"),
                                                  ("27pt-10-random3", lambda: mg.stencil3d_27pt(10), 3, "random")])
def test_every_part_byte_identical(name, gen, nparts, kind, ab):
    from acg_b200 import dist as abdist
    n, r, c, v = gen()
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v)
    if kind == "block":
        N = round(n ** (1 / 3))
        rowparts = abdist.block_partition(N, N, N, 2, 2, 2)
    elif kind == "contiguous":
        rowparts = (np.arange(n) * nparts // n).astype(np.int32)
    else:
        rowparts = np.random.default_rng(5).integers(0, nparts, n).astype(np.int32)
    hosts = A.partition(nparts, rowparts)
    devs = A.partition(nparts, rowparts)
    for h, d in zip(hosts, devs):
        h.dsymv_init(0.25)
        d.dsymv_init_cuda(0.25)
        assert h.c.nghostrows > 0 and h.c.onpnzs > 0
        _same(h, d)


@pytest.mark.parametrize("path", GOLDEN, ids=[p.split("/")[-1] for p in GOLDEN])
def test_golden_full_storage(path, ab):
    g = load_golden(path)
    dev = ab.SymCsrMatrix.init_real_double(int(g["n"]), g["rows"], g["cols"], g["vals"]).dsymv_init_cuda(0.0)
    assert np.array_equal(dev.frowptr, g["frowptr"]) and np.array_equal(dev.fcolidx, g["fcolidx"])
    assert dev.fa.tobytes() == g["fa"].tobytes()


def test_one_based_indices(ab):
    n, r, c, v = mg.stencil3d_27pt(9)
    host = ab.SymCsrMatrix.init_real_double(n, r + 1, c + 1, v, idxbase=1).dsymv_init(0.5)
    dev = ab.SymCsrMatrix.init_real_double(n, r + 1, c + 1, v, idxbase=1).dsymv_init_cuda(0.5)
    _same(host, dev)


@pytest.mark.parametrize("name,gen", [("27pt-aniso", lambda: mg.stencil3d_27pt(9, 40, 17)), ("rmat", lambda: mg.rmat_spd(30000, 600000, seed=8))])
def test_solver_without_full_storage(name, gen, ab, oracle):
    """acgsolvercuda_init on a matrix on which acgsymcsrmatrix_dsymv_init was never called."""
    n, r, c, v = gen()
    packed = ab.SymCsrMatrix.init_real_double(n, r, c, v)
    assert len(packed.frowptr) == 0
    cg = ab.SolverCuda(packed)
    ref_m = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    csr = (ref_m.frowptr.copy(), ref_m.fcolidx.copy(), ref_m.fa.copy())
    x = np.random.default_rng(1).standard_normal(n)
    y, _ = cg.spmv(x)
    want = oracle.dsymv(csr, 1.0, x, 0.0, np.zeros(n))
    scale = oracle.dsymv((csr[0], csr[1], np.abs(csr[2])), 1.0, np.abs(x), 0.0, np.zeros(n))
    assert np.all(np.abs(y - want) <= 1e-13 * scale + 1e-300)
    b = ref_m.vector(); b.x[:] = 1.0 + 0.5 * np.sin(0.37 * np.arange(n))
    # the power-law matrix is ill conditioned (hubs): rounding differences between the device's and the
    # oracle's summation orders grow with the iteration count -- compare early
    its, xtol = (60, 1e-9) if name != "rmat" else (12, 1e-7)
    for method, oname in (("solvempi", "cg"), ("solve_pipelined", "cg_pipelined")):
        xs = ref_m.vector()
        ref = getattr(oracle, oname)(csr, b.x, maxits=its, rtol=0.0)
        assert getattr(cg, method)(b, xs, maxits=its) == 0 and cg.c.niterations == its
        assert np.abs(xs.x - ref["x"]).max() <= xtol * np.abs(ref["x"]).max()
    cg.free()
