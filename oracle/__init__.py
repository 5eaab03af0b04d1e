"""ctypes front-end for the CPU checker libraries under oracle/.

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; the product package
(acg_b200) never imports this.

``Oracle``  -> oracle/liboracle.so       (our restatement, cg_oracle.c)
``Ref``     -> oracle/_ref/libacgref.so  (the unmodified reference sources +
                                          ref_shim.c), present when built in a
                                          container that has /root/reference
                                          (the built .so travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """(Re)build liboracle.so and, if the reference tree is present, _ref/."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Result(C.Structure):
    _fields_ = [("status", C.c_int), ("niterations", C.c_int),
                ("bnrm2", C.c_double), ("r0nrm2", C.c_double), ("rnrm2", C.c_double)]


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.oracle_full_csr.restype = C.c_int64
        L.oracle_full_csr.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_double, _i64, _i32, _f64]
        L.oracle_dsymv.restype = None
        L.oracle_dsymv.argtypes = [C.c_int, _i64, _i32, _f64, C.c_double, _f64, C.c_double, _f64]
        L.oracle_ddot.restype = C.c_double
        L.oracle_ddot.argtypes = [C.c_int, _f64, _f64]
        L.oracle_dnrm2sqr.restype = C.c_double
        L.oracle_dnrm2sqr.argtypes = [C.c_int, _f64]
        L.oracle_daxpy.restype = None
        L.oracle_daxpy.argtypes = [C.c_int, C.c_double, _f64, _f64]
        L.oracle_daypx.restype = None
        L.oracle_daypx.argtypes = [C.c_int, C.c_double, _f64, _f64]
        for f in (L.oracle_cg, L.oracle_cg_pipelined):
            f.restype = None
            f.argtypes = [C.c_int, _i64, _i32, _f64, _f64, _f64, C.c_int, C.c_double, C.c_double,
                          C.POINTER(_Result), C.c_void_p]
        L.oracle_num_threads.restype = C.c_int

    def num_threads(self) -> int:
        return self.lib.oracle_num_threads()

    def full_csr(self, n, rows, cols, vals, eps=0.0):
        nnz = len(vals)
        rp = np.zeros(n + 1, np.int64)
        ci = np.zeros(max(2 * nnz, 1), np.int32)
        va = np.zeros(max(2 * nnz, 1), np.float64)
        f = self.lib.oracle_full_csr(n, nnz, rows, cols, vals, eps, rp, ci, va)
        assert f >= 0
        return rp, ci[:f].copy(), va[:f].copy()

    def dsymv(self, csr, alpha, x, beta, y):
        rp, ci, va = csr
        y = np.array(y, dtype=np.float64, copy=True)
        self.lib.oracle_dsymv(len(rp) - 1, rp, ci, va, alpha, np.ascontiguousarray(x, np.float64), beta, y)
        return y

    def ddot(self, x, y):
        return self.lib.oracle_ddot(len(x), x, y)

    def dnrm2sqr(self, x):
        return self.lib.oracle_dnrm2sqr(len(x), x)

    def _solve(self, fn, csr, b, x0, maxits, atol, rtol, history):
        rp, ci, va = csr
        n = len(rp) - 1
        x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        hist = np.full(maxits + 1, np.nan) if history else None
        res = _Result()
        fn(n, rp, ci, va, np.ascontiguousarray(b, np.float64), x, maxits, atol, rtol, C.byref(res),
           hist.ctypes.data if history else None)
        out = dict(status=res.status, niterations=res.niterations, bnrm2=res.bnrm2,
                   r0nrm2=res.r0nrm2, rnrm2=res.rnrm2, x=x)
        if history:
            out["rnrm2hist"] = hist
        return out

    def cg(self, csr, b, x0=None, maxits=100, atol=0.0, rtol=0.0, history=False):
        return self._solve(self.lib.oracle_cg, csr, b, x0, maxits, atol, rtol, history)

    def cg_pipelined(self, csr, b, x0=None, maxits=100, atol=0.0, rtol=0.0, history=False):
        return self._solve(self.lib.oracle_cg_pipelined, csr, b, x0, maxits, atol, rtol, history)


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libacgref.so"))


class Ref:
    """The reference's own CPU path (acg/cg.c, acg/symcsrmatrix.c, acg/vector.c)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libacgref.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.ref_full_csr.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_double, _i64, _i32, _f64,
                                   C.POINTER(C.c_int64)]
        L.ref_dsymv.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_double, _f64, C.c_double, _f64]
        L.ref_ddot.restype = C.c_double
        L.ref_ddot.argtypes = [C.c_int, _f64, _f64]
        L.ref_dnrm2sqr.restype = C.c_double
        L.ref_dnrm2sqr.argtypes = [C.c_int, _f64]
        L.ref_daxpy.restype = None
        L.ref_daxpy.argtypes = [C.c_int, C.c_double, _f64, _f64]
        L.ref_daypx.restype = None
        L.ref_daypx.argtypes = [C.c_int, C.c_double, _f64, _f64]
        L.ref_cg.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_double, _f64, _f64, C.c_int,
                             C.c_double, C.c_double, _f64]
        L.ref_num_threads.restype = C.c_int
        L.ref_setup.restype = C.c_void_p
        L.ref_setup.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_double]
        L.ref_solve.argtypes = [C.c_void_p, _f64, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, _f64]
        L.ref_free.restype = None
        L.ref_free.argtypes = [C.c_void_p]
        if hasattr(L, "ref_place"):
            L.ref_place.argtypes = [C.c_void_p]
        vp = C.c_void_p
        L.ref_partition_part.argtypes = [C.c_int, C.c_int64, _i32, _i32, _f64, C.c_int, _i32, C.c_int,
                                         _i64] + [vp] * 13

    def num_threads(self) -> int:
        return self.lib.ref_num_threads()

    def full_csr(self, n, rows, cols, vals, eps=0.0):
        nnz = len(vals)
        rp = np.zeros(n + 1, np.int64)
        ci = np.zeros(max(2 * nnz, 1), np.int32)
        va = np.zeros(max(2 * nnz, 1), np.float64)
        f = C.c_int64(0)
        err = self.lib.ref_full_csr(n, nnz, rows, cols, vals, eps, rp, ci, va, C.byref(f))
        assert err == 0, err
        return rp, ci[:f.value].copy(), va[:f.value].copy()

    def dsymv(self, n, rows, cols, vals, alpha, x, beta, y):
        y = np.array(y, dtype=np.float64, copy=True)
        err = self.lib.ref_dsymv(n, len(vals), rows, cols, vals, alpha, np.ascontiguousarray(x, np.float64), beta, y)
        assert err == 0, err
        return y

    def ddot(self, x, y):
        return self.lib.ref_ddot(len(x), x, y)

    def dnrm2sqr(self, x):
        return self.lib.ref_dnrm2sqr(len(x), x)

    def cg(self, n, rows, cols, vals, b, x0=None, maxits=100, atol=0.0, rtol=0.0, eps=0.0):
        x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        out = np.zeros(6)
        status = self.lib.ref_cg(n, len(vals), rows, cols, vals, eps, np.ascontiguousarray(b, np.float64), x,
                                 maxits, atol, rtol, out)
        return dict(status=status, niterations=int(out[0]), bnrm2=out[1], r0nrm2=out[2], rnrm2=out[3],
                    tsolve=out[4], tgemv=out[5], x=x)

    def setup(self, n, rows, cols, vals, eps=0.0):
        """Persistent problem handle (reference setup path run once)."""
        h = self.lib.ref_setup(n, len(vals), rows, cols, vals, eps)
        assert h, "ref_setup failed"
        return h

    def solve(self, handle, b, maxits=100, atol=0.0, rtol=0.0, want_x=False):
        out = np.zeros(6)
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(len(b)) if want_x else None
        status = self.lib.ref_solve(handle, b, None, x.ctypes.data if want_x else None, maxits, atol, rtol, out)
        return dict(status=status, niterations=int(out[0]), bnrm2=out[1], r0nrm2=out[2], rnrm2=out[3],
                    tsolve=out[4], tgemv=out[5], x=x)

    def place(self, handle) -> int:
        """First-touch the full-storage arrays by the threads that stream them (ref_shim.c, ref_place)."""
        return int(self.lib.ref_place(handle)) if hasattr(self.lib, "ref_place") else 0

    def free(self, handle):
        self.lib.ref_free(handle)

    def partition_part(self, n, rows, cols, vals, nparts, rowparts, p):
        """Local structure of part p as the reference builds it (see ref_shim.c)."""
        sz = np.zeros(11, np.int64)
        nul = [None] * 13
        rowparts = np.ascontiguousarray(rowparts, np.int32)
        err = self.lib.ref_partition_part(n, len(vals), rows, cols, vals, nparts, rowparts, p, sz, *nul)
        assert err == 0, err
        nprows, nowned, ninner, nborder, nghost, nrecip, sendsize, nsend, recvsize, fnnz, onnz = (int(v) for v in sz)
        arrs = dict(
            nzrows=np.zeros(nprows, np.int32),
            recipients=np.zeros(nrecip, np.int32), sendcounts=np.zeros(nrecip, np.int32),
            sendbufidx=np.zeros(sendsize, np.int32),
            senders=np.zeros(nsend, np.int32), recvcounts=np.zeros(nsend, np.int32),
            recvbufidx=np.zeros(recvsize, np.int32),
            frowptr=np.zeros(nprows + 1, np.int64), fcolidx=np.zeros(fnnz, np.int32), fa=np.zeros(fnnz),
            orowptr=np.zeros(nborder + nghost + 1, np.int64), ocolidx=np.zeros(onnz, np.int32), oa=np.zeros(onnz))
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs.values()]
        err = self.lib.ref_partition_part(n, len(vals), rows, cols, vals, nparts, rowparts, p, sz, *ptrs)
        assert err == 0, err
        arrs.update(nprows=nprows, nownedrows=nowned, ninnerrows=ninner, nborderrows=nborder, nghostrows=nghost)
        return arrs
