"""SpMV tile-plan sweep on the GPU box (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import acg_b200 as ab
from acg_b200 import matgen as mg

N = int(os.environ.get("N", "224"))
kind = os.environ.get("KIND", "27")
n, r, c, v = (mg.stencil3d_27pt(N) if kind == "27" else mg.laplace3d_7pt(N))
A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
del r, c, v
nnz = A.c.fnpnzs
xin = np.random.default_rng(1).standard_normal(n)
yref = None
cfgs = []
for line in open(sys.argv[1]):
    line = line.split("#")[0].strip()
    if line:
        cfgs.append(tuple(int(t) for t in line.split()))
for (lanes, threads, unroll, rows, nnzcap, stages) in cfgs:
    for k, val in (("spmv_lanes", lanes), ("spmv_threads", threads), ("spmv_unroll", unroll), ("spmv_rows_cap", rows), ("spmv_nnz_cap", nnzcap), ("spmv_stages", stages)):
        ab.set_option(k, val)
    try:
        cg = ab.SolverCuda(A)
    except Exception as e:
        print(f"G={lanes} T={threads} U={unroll} rows={rows} nnzcap={nnzcap} st={stages}: init failed {e}")
        continue
    y, ms = cg.spmv(xin, nrep=10)
    if yref is None: yref = y
    err = np.abs(y - yref).max() / np.abs(yref).max()
    inf = cg.info()
    print(f"G={lanes} T={threads} U={unroll} rows={rows} nnzcap={nnzcap} st={stages}: {ms:.4f} ms {16*nnz/ms/1e6:.0f} GB/s(16nnz) {(12*nnz+20*n)/ms/1e6:.0f} GB/s(min) grid={inf['spmv_grid']} ({inf['spmv_grid']//148}/SM) smem={inf['spmv_smem_bytes']} err={err:.1e}", flush=True)
    cg.free()
