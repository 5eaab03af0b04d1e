"""TEST INFRASTRUCTURE ONLY: set-up and solves with the k-th "device" allocation failing
(HOSTSIM_FAIL_MALLOC_AT, tests/hostsim/cuda_mock.c): the library must return ACG_ERR_CUDA and
release what it had built -- no crash, no use of freed memory (see tools/asan_hostsim.sh)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import acg_b200.api as api                           # noqa: E402
api._LIBPATH = os.environ.get("ACGB200_TEST_LIB") or os.path.join(HERE, "libacgb200_hostsim.so")
import acg_b200 as ab                                # noqa: E402
from acg_b200 import matgen as mg                    # noqa: E402

n, r, c, v = mg.stencil3d_27pt(7)
A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
try:
    cg = ab.SolverCuda(A)                             # pattern slices + tiles (default)
    b = A.vector(); b.x[:] = 1.0; x = A.vector()
    cg.solvempi(b, x, maxits=8, warmup=1)
    cg.solve_pipelined(b, x, maxits=8, warmup=1)
    cg.free()
    cg = ab.SolverCuda(ab.SymCsrMatrix.init_real_double(n, r, c, v))      # no full storage: device-side expansion
    cg.solve_pipelined(b, x, maxits=8)
    cg.free()
    A2 = ab.SymCsrMatrix.init_real_double(n, r, c, v)                      # acgsymcsrmatrix_dsymv_init_cuda
    try:
        A2.dsymv_init_cuda(0.0)
    except ab.AcgError:
        # a failed expansion leaves no half-filled full storage behind
        assert not A2.c.frowptr and not A2.c.fcolidx and not A2.c.fa and not A2.c.orowptr and A2.c.fnpnzs == 0
        raise
    assert A2.c.fnpnzs == A.c.fnpnzs
    print("ok")
except ab.AcgError as e:
    print("error", e.code)
