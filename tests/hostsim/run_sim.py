"""TEST INFRASTRUCTURE ONLY: drive the solver's host code on the device stand-in
(tests/hostsim/libacgb200_hostsim.so) and compare with the oracle.  Runs in its own
process so that the binding loads the stand-in instead of the product library; prints
one JSON object.  Used by tests/test_hostsim.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np                                   # noqa: E402

import acg_b200.api as api                           # noqa: E402
api._LIBPATH = os.path.join(HERE, "libacgb200_hostsim.so")      # before the first lib() call
import acg_b200 as ab                                # noqa: E402
from acg_b200 import matgen as mg                    # noqa: E402
from oracle import Oracle                            # noqa: E402


def case(name):
    return {"27pt": lambda: mg.stencil3d_27pt(9, 8, 10), "7pt": lambda: mg.laplace3d_7pt(11),
            "rmat": lambda: mg.rmat_spd(30000, 600000, seed=8), "n3": lambda: mg.poisson1d_3pt(3)}[name]()


def blas1_blocks():
    """acg/cg-kernels-cuda.h:45-97 through the C-ABI; on the stand-in "device" pointers are host addresses."""
    import ctypes as C
    L = ab.lib()
    rng = np.random.default_rng(2)
    n = 1001
    ptr = lambda a: a.ctypes.data                                    # noqa: E731
    out = {}
    rr, pap, rrp = np.array([3.5]), np.array([1.25]), np.array([7.0])
    al, mal, be = np.zeros(1), np.zeros(1), np.zeros(1)
    assert L.acgsolvercuda_alpha(ptr(al), ptr(mal), ptr(rr), ptr(pap)) == 0
    assert L.acgsolvercuda_beta(ptr(be), ptr(rr), ptr(rrp)) == 0
    out["scalars"] = bool(al[0] == 2.8 and mal[0] == -2.8 and be[0] == 0.5)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    y1 = y.copy(); assert L.acgsolvercuda_daxpy_alpha(n, ptr(rr), ptr(pap), ptr(x), ptr(y1)) == 0
    y2 = y.copy(); assert L.acgsolvercuda_daxpy_minus_alpha(n, ptr(rr), ptr(pap), ptr(x), ptr(y2)) == 0
    y3 = y.copy(); assert L.acgsolvercuda_daypx_beta(n, ptr(rr), ptr(rrp), ptr(y3), ptr(x)) == 0
    out["axpy"] = bool(np.allclose(y1, y + 2.8 * x, rtol=1e-15, atol=1e-15) and np.allclose(y2, y - 2.8 * x, rtol=1e-15, atol=1e-15)
                       and np.allclose(y3, 0.5 * y + x, rtol=1e-15, atol=1e-15))
    g, gp, d, ap = np.array([2.0]), np.array([4.0]), np.array([3.0]), np.array([0.5])
    v = {k: rng.standard_normal(n) for k in "qprtxzw"}
    w0 = {k: a.copy() for k, a in v.items()}
    assert L.acgsolvercuda_pipelined_daxpy_fused(n, ptr(g), ptr(gp), ptr(d), ptr(v["q"]), ptr(v["p"]), ptr(v["r"]), ptr(v["t"]),
                                                 ptr(v["x"]), ptr(v["z"]), ptr(v["w"]), ptr(ap), None) == 0
    beta = 2.0 / 4.0; alpha = 2.0 / (3.0 - beta * 2.0 / 0.5)
    z = w0["q"] + beta * w0["z"]; t = w0["w"] + beta * w0["t"]; p = w0["r"] + beta * w0["p"]
    ok = (np.allclose(v["z"], z) and np.allclose(v["t"], t) and np.allclose(v["p"], p) and np.allclose(v["x"], w0["x"] + alpha * p)
          and np.allclose(v["r"], w0["r"] - alpha * t) and np.allclose(v["w"], w0["w"] - alpha * z))
    out["pipelined"] = bool(ok and gp[0] == 2.0 and ap[0] == alpha)
    m1, p1, z0 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert L.acgsolvercuda_init_constants(C.byref(m1), C.byref(p1), C.byref(z0)) == 0
    vals = [C.cast(q.value, C.POINTER(C.c_double))[0] for q in (m1, p1, z0)]
    out["constants"] = vals == [-1.0, 1.0, 0.0]
    return out


def main():
    spec = json.loads(sys.argv[1])
    for k, v in spec.get("options", {}).items():
        ab.set_option(k, v)
    O = Oracle()
    n, r, c, v = case(spec["matrix"])
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    csr = (A.frowptr.copy(), A.fcolidx.copy(), A.fa.copy())
    # "no_full_storage": the solver gets a matrix on which acgsymcsrmatrix_dsymv_init was never called
    # and expands the packed triangle on the (stand-in) device itself
    cg = ab.SolverCuda(ab.SymCsrMatrix.init_real_double(n, r, c, v) if spec.get("no_full_storage") else A)
    inf = cg.info()
    b = A.vector(); b.x[:] = np.random.default_rng(5).standard_normal(n)
    out = {"merge_tiles": inf["spmv_merge_tiles"], "merge_rows": inf["spmv_merge_rows"], "merge_split": inf["spmv_merge_split"],
           "slices": inf["spmv_slices"], "slice_rows": inf["spmv_slice_rows"], "nlong": inf["spmv_nlong"], "nmedium": inf["spmv_nmedium"], "ntiles": inf["spmv_ntiles"], "runs": []}
    y, _ = cg.spmv(b.x)
    want = O.dsymv(csr, 1.0, b.x, 0.0, np.zeros(n))
    out["spmv_err"] = float(np.abs(y - want).max() / np.abs(want).max())
    for run in spec["runs"]:
        method = run["method"]
        orc = O.cg_pipelined if "pipelined" in method else O.cg
        ref = orc(csr, b.x, maxits=run["maxits"], rtol=run.get("rtol", 0.0))
        x = A.vector()
        code = getattr(cg, method)(b, x, maxits=run["maxits"], residualrtol=run.get("rtol", 0.0), warmup=run.get("warmup", 0))
        out["runs"].append({
            "method": method, "maxits": run["maxits"], "code": int(code), "ref_code": int(ref["status"]),
            "its": int(cg.c.niterations), "ref_its": int(ref["niterations"]),
            "rnrm2": float(cg.c.rnrm2), "ref_rnrm2": float(ref["rnrm2"]),
            "r0nrm2": float(cg.c.r0nrm2), "ref_r0nrm2": float(ref["r0nrm2"]), "bnrm2": float(cg.c.bnrm2),
            "xerr": float(np.abs(x.x - ref["x"]).max() / max(np.abs(ref["x"]).max(), 1e-300)),
            "launches": int(cg.info()["last_launches"]), "nsolves": int(cg.c.nsolves),
            "spmv_count": int(cg.info()["last_spmv_count"]), "ngemv": int(cg.c.ngemv), "total_its": int(cg.c.ntotaliterations)})
    out["report_ok"] = "total solver time:" in cg.report()
    if spec.get("blas1_blocks"):
        out["blas1_blocks"] = blas1_blocks()
    # interface behaviour that needs no kernel at all
    errs = {}
    try:
        cg.solvempi(b, A.vector(), maxits=5, diffatol=1e-3)
    except ab.AcgError as e:
        errs["diffatol"] = e.code
    short = ab.Vector(max(n - 1, 1))
    try:
        cg.solvempi(short, A.vector(), maxits=5)
    except ab.AcgError as e:
        errs["short_b"] = e.code
    out["errors"] = errs
    out["nsolves_after_errors"] = int(cg.c.nsolves)
    cg.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
