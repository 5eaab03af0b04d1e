"""Randomised checks of the host data structures (CPU only):

    python tools/fuzz_host.py partition SEED TRIALS   # partition + full storage + tile plans: the parts' products = the global product
    python tools/fuzz_host.py ingest SEED TRIALS      # acgb200_mtx_read_part = acgsymcsrmatrix_partition, array for array

Set ACGB200_TEST_LIB to another build of the library's sources (e.g. the AddressSanitizer build of
tools/asan_hostsim.sh, together with LD_PRELOAD=libasan.so) to run them under a sanitizer.  Both were clean
at the end of round 1 (150 + 60 trials under ASan/UBSan)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                  # noqa: E402
import acg_b200.api as api                          # noqa: E402
if os.environ.get("ACGB200_TEST_LIB"):
    api._LIBPATH = os.environ["ACGB200_TEST_LIB"]
import acg_b200 as ab                               # noqa: E402
from acg_b200 import matgen as mg, mtxio            # noqa: E402
from oracle import Oracle                           # noqa: E402


def fuzz_partition(seed, trials):
    O = Oracle()
    rng = np.random.default_rng(seed)
    bad = 0
    for trial in range(trials):
        n = int(rng.integers(2, 400))
        kind = rng.integers(0, 3)
        if kind == 0:
            nn, r, c, v = mg.rmat_spd(n, int(rng.integers(n, 8 * n)), seed=int(rng.integers(1 << 30)))
        elif kind == 1:
            nn, r, c, v = mg.random_spd(n, float(rng.uniform(0.01, 0.3)), seed=int(rng.integers(1 << 30)))
        else:
            a, b_, c_ = [int(x) for x in rng.integers(2, 8, 3)]
            nn, r, c, v = mg.stencil3d_27pt(a, b_, c_)
        nparts = int(rng.integers(1, 8))
        rowparts = rng.integers(0, nparts, nn).astype(np.int32)
        A = ab.SymCsrMatrix.init_real_double(nn, r, c, v)
        parts = A.partition(nparts, rowparts)
        csr = O.full_csr(nn, r, c, v)
        x = rng.standard_normal(nn)
        want = O.dsymv(csr, 1.0, x, 0.0, np.zeros(nn))
        got = np.zeros(nn)
        owned_total = 0
        for p, m in enumerate(parts):
            m.dsymv_init(0.0)
            no, npn = m.c.nownedrows, m.c.nprows
            gi = m.nzrows[:npn]
            xl = x[gi]
            import scipy.sparse as sp
            rp = m.frowptr[:no + 1]
            F = sp.csr_matrix((m.fa[:rp[no]], m.fcolidx[:rp[no]], rp), shape=(no, npn))
            y = F @ xl
            if m.c.onpnzs > 0:
                nb, off = m.c.nborderrows, m.c.borderrowoffset
                orp = m.orowptr[:nb + 1]
                Ob = sp.csr_matrix((m.oa[:orp[nb]], m.ocolidx[:orp[nb]] + off, orp), shape=(nb, npn))
                y[off:off + nb] += Ob @ xl
            got[gi[:no]] = y
            owned_total += no
            h = m.halo()
            pl = ab.spmv_plan_host(m.frowptr[:no + 1].copy())
            assert sum(t[1] for t in pl["tiles"]) + len(pl["longrows"]) + pl["nmedium"] == no
        assert owned_total == nn
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
        if err > 1e-12:
            bad += 1; print("MISMATCH", trial, nn, nparts, err)
    print("trials done, bad =", bad)


def fuzz_ingest(seed, trials):
    rng = np.random.default_rng(seed)
    d = tempfile.mkdtemp()
    for trial in range(trials):
        n = int(rng.integers(2, 300))
        nn, r, c, v = mg.rmat_spd(n, int(rng.integers(n, 6 * n)), seed=int(rng.integers(1 << 30)))
        v = v * (1 + 1e-3 * np.arange(len(v)))
        if rng.integers(0, 2):                       # lower-triangle file
            r, c = c, r
        path = os.path.join(d, f"A{trial}.mtx")
        mtxio.write_symmetric(path, nn, r, c, v, binary=True)
        nparts = int(rng.integers(1, 6))
        rowparts = rng.integers(0, nparts, nn).astype(np.int32)
        want = ab.SymCsrMatrix.init_real_double(nn, r, c, v).partition(nparts, rowparts)
        for p in range(nparts):
            got = ab.SymCsrMatrix.read_mtx_part(path, nparts, rowparts, p)
            w = want[p]
            for k in ("nprows", "npnzs", "nownedrows", "ninnerrows", "nborderrows", "nghostrows"):
                assert getattr(got.c, k) == getattr(w.c, k), (trial, p, k)
            assert np.array_equal(got.rowptr, w.rowptr) and np.array_equal(got.colidx, w.colidx) and np.array_equal(got.a, w.a)
            got.dsymv_init(0.0); w.dsymv_init(0.0)
            for k in ("frowptr", "fcolidx", "fa", "orowptr", "ocolidx", "oa"):
                assert np.array_equal(getattr(got, k), getattr(w, k)), (trial, p, k)
            got.free()
        os.remove(path)
    print("mtx fuzz ok")


if __name__ == "__main__":
    {"partition": fuzz_partition, "ingest": fuzz_ingest}[sys.argv[1]](int(sys.argv[2]), int(sys.argv[3]))
