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


def main():
    spec = json.loads(sys.argv[1])
    for k, v in spec.get("options", {}).items():
        ab.set_option(k, v)
    O = Oracle()
    n, r, c, v = case(spec["matrix"])
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    csr = (A.frowptr.copy(), A.fcolidx.copy(), A.fa.copy())
    cg = ab.SolverCuda(A)
    inf = cg.info()
    b = A.vector(); b.x[:] = np.random.default_rng(5).standard_normal(n)
    out = {"nlong": inf["spmv_nlong"], "nmedium": inf["spmv_nmedium"], "ntiles": inf["spmv_ntiles"], "compressed_tiles": inf["spmv_compressed_tiles"], "runs": []}
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
