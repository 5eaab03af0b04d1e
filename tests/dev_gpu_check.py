"""Development check run on the GPU box: parity on small cases + SpMV timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import acg_b200 as ab
from acg_b200 import matgen as mg
from oracle import Oracle

O = Oracle()

def parity(name, n, r, c, v, maxits=300, rtol=1e-10, rhs="ones"):
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    csr = (A.frowptr.copy(), A.fcolidx.copy(), A.fa.copy())
    b = A.vector()
    b.x[:] = 1.0 if rhs == "ones" else np.random.default_rng(5).standard_normal(n)
    cg = ab.SolverCuda(A)
    xin = np.random.default_rng(1).standard_normal(n)
    y, _ = cg.spmv(xin)
    yo = O.dsymv(csr, 1.0, xin, 0.0, np.zeros(n))
    print(f"{name}: n={n} spmv max rel err {np.abs(y-yo).max()/np.abs(yo).max():.2e}", cg.info())
    for meth, orc in (("solvempi", O.cg), ("solve_pipelined", O.cg_pipelined)):
        x = A.vector()
        code = getattr(cg, meth)(b, x, maxits=maxits, residualrtol=rtol, warmup=2)
        ref = orc(csr, b.x, maxits=maxits, rtol=rtol)
        err = np.abs(x.x - ref["x"]).max() / max(np.abs(ref["x"]).max(), 1e-300)
        print(f"  {meth}: status {code}/{ref['status']} its {cg.c.niterations}/{ref['niterations']} rnrm2 {cg.c.rnrm2:.6e}/{ref['rnrm2']:.6e} r0 {cg.c.r0nrm2:.6e}/{ref['r0nrm2']:.6e} bnrm2 {cg.c.bnrm2:.6e} xerr {err:.2e}")
    cg.free(); A.free()

if "parity" in sys.argv or len(sys.argv) == 1:
    parity("27pt-12", *mg.stencil3d_27pt(12))
    parity("27pt-33", *mg.stencil3d_27pt(33))
    parity("7pt-31", *mg.laplace3d_7pt(31, 17, 23))
    parity("1d5pt", *mg.poisson1d_5pt(100000), maxits=200, rtol=0)
    parity("1d3pt", *mg.poisson1d_3pt(1000), maxits=2000)
    parity("rand", *mg.random_spd(300, 0.3, 2), rhs="rand")
    parity("rmat", *mg.rmat_spd(20000, 400000), rhs="rand")
    parity("tiny1", *mg.poisson1d_3pt(1))
    parity("tiny5", *mg.poisson1d_3pt(5))

if "time" in sys.argv:
    N = int(os.environ.get("N", "128"))
    t0 = time.time(); n, r, c, v = mg.stencil3d_27pt(N); t1 = time.time()
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v); t2 = time.time()
    A.dsymv_init(0.0); t3 = time.time()
    print(f"gen {t1-t0:.1f}s init {t2-t1:.1f}s dsymv_init {t3-t2:.1f}s n={n} nnz={A.c.fnpnzs}")
    del r, c, v
    nnz = A.c.fnpnzs
    xin = np.random.default_rng(1).standard_normal(n)
    for lanes, nnzcap, rowscap, stages in [(0,0,0,0), (1,3456,128,2), (1,3456,128,3), (1,3456,128,4), (1,6912,256,2), (2,3456,128,3), (4,3456,128,3), (1,1728,64,4), (1,1728,64,6), (8,4096,128,3)]:
        ab.set_option("spmv_lanes", lanes); ab.set_option("spmv_nnz_cap", nnzcap); ab.set_option("spmv_rows_cap", rowscap); ab.set_option("spmv_stages", stages)
        t4 = time.time(); cg = ab.SolverCuda(A); t5 = time.time()
        y, ms = cg.spmv(xin, nrep=20)
        inf = cg.info()
        print(f"cfg lanes={lanes} nnzcap={nnzcap} rows={rowscap} st={stages}: init {t5-t4:.1f}s spmv {ms:.4f} ms  {16*nnz/ms/1e6:.0f} GB/s(16nnz) {(12*nnz+20*n)/ms/1e6:.0f} GB/s(actual) grid={inf['spmv_grid']} smem={inf['spmv_smem_bytes']} tiles={inf['spmv_ntiles']}")
        cg.free()
    ab.set_option("spmv_lanes", 0); ab.set_option("spmv_nnz_cap", 0); ab.set_option("spmv_rows_cap", 0); ab.set_option("spmv_stages", 0)
    ab.set_option("profile", 1)
    cg = ab.SolverCuda(A)
    b = A.vector(); b.x[:] = 1.0
    for meth in ("solvempi", "solve_pipelined"):
        x = A.vector()
        t6 = time.time(); code = getattr(cg, meth)(b, x, maxits=100, warmup=3); t7 = time.time()
        inf = cg.info()
        print(f"{meth}: total call {t7-t6:.3f}s tsolve {cg.c.tsolve:.4f} its {cg.c.niterations} -> {cg.c.niterations/cg.c.tsolve:.1f} it/s; spmv {inf['last_spmv_ms']/max(inf['last_spmv_count'],1):.4f} ms avg over {inf['last_spmv_count']}; rnrm2/r0 {cg.c.rnrm2/cg.c.r0nrm2:.3e}")
        cg.c.tsolve = 0
    cg.free()
