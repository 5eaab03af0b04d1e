"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container, where /root/reference exists:
    python tools/make_golden.py
Each fixture holds a small SPD system (upper COO + rhs), the reference's
full-storage CSR, one dsymv result, and the result of acgsolver_solve
(acg/cg.c:198): x, iteration count, ||b||, ||r0||, ||r||, status.  Nothing is
computed by this repository's own code here.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acg_b200 import matgen as mg
from oracle import Ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
R = Ref()
rng = np.random.default_rng(20260921)

cases = {
    # name: (generator output, rhs kind, maxits, rtol)
    "poisson1d_3pt_n100": (mg.poisson1d_3pt(100), "ones", 200, 1e-10),          # KAT-1: 50 iterations, x_i=(i+1)(n-i)/2
    "poisson1d_5pt_n400_fixedits": (mg.poisson1d_5pt(400), "ones", 40, 0.0),    # config 1 shape, fixed iteration count
    "laplace7_5x6x7": (mg.laplace3d_7pt(5, 6, 7), "ones", 100, 1e-9),
    "stencil27_6": (mg.stencil3d_27pt(6), "rand", 100, 1e-9),
    "stencil27_9_notconverged": (mg.stencil3d_27pt(9), "ones", 5, 1e-12),      # returns ACG_ERR_NOT_CONVERGED
    "random_spd_60": (mg.random_spd(60, 0.2, 7), "rand", 100, 1e-11),
    "rmat_300": (mg.rmat_spd(300, 2500, seed=3), "rand", 100, 1e-10),
}
for name, ((n, r, c, v), rhs, maxits, rtol) in cases.items():
    b = np.ones(n) if rhs == "ones" else rng.standard_normal(n)
    xs = rng.standard_normal(n)
    frowptr, fcolidx, fa = R.full_csr(n, r, c, v)
    y = R.dsymv(n, r, c, v, 1.0, xs, 0.0, np.zeros(n))
    y2 = R.dsymv(n, r, c, v, -1.0, xs, 1.0, b)
    res = R.cg(n, r, c, v, b, maxits=maxits, rtol=rtol)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, rows=r, cols=c, vals=v, b=b, xs=xs,
                        frowptr=frowptr, fcolidx=fcolidx, fa=fa, y=y, y2=y2, maxits=maxits, rtol=rtol,
                        x=res["x"], niterations=res["niterations"], status=res["status"],
                        bnrm2=res["bnrm2"], r0nrm2=res["r0nrm2"], rnrm2=res["rnrm2"])
    print(f"{name}: n={n} nnz_upper={len(v)} its={res['niterations']} status={res['status']} rnrm2={res['rnrm2']:.3e}")
