"""Matrix Market ingest (acg_b200/csrc/mtxfile.c): whole-matrix read in both
encodings and the per-rank read that replaces read-on-root + scatter
(cuda/acg-cuda.c:1297-1304, :1516-1782).  CPU only."""
import os

import numpy as np
import pytest

from acg_b200 import dist as abdist
from acg_b200 import matgen as mg
from acg_b200 import mtxio

FIELDS = ("nrows", "nprows", "nnzs", "npnzs", "nownedrows", "ninnerrows", "nborderrows", "borderrowoffset",
          "nghostrows", "ghostrowoffset", "ninnernzs", "ninterfacenzs")


def _same(m, w):
    for k in FIELDS:
        assert getattr(m.c, k) == getattr(w.c, k), k
    assert np.array_equal(m.rowptr, w.rowptr) and np.array_equal(m.colidx, w.colidx) and np.array_equal(m.a, w.a)
    if w.c.nghostrows or w.c.nborderrows:
        assert np.array_equal(m.nzrows, w.nzrows)
    hm, hw = m.halo(), w.halo()
    for k in hm:
        assert np.array_equal(hm[k], hw[k]), k
    m.dsymv_init(0.0); w.dsymv_init(0.0)
    for k in ("frowptr", "fcolidx", "fa", "orowptr", "ocolidx", "oa"):
        assert np.array_equal(getattr(m, k), getattr(w, k)), k


@pytest.mark.parametrize("binary", [True, False], ids=["binary", "text"])
def test_read_whole_matrix(binary, ab, tmp_path):
    n, r, c, v = mg.rmat_spd(600, 4000, seed=4)
    v = v * (1 + 0.001 * np.arange(len(v)))               # distinct values: any mix-up shows
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=binary)
    inf = ab.mtx_info(path)
    assert (inf["nrows"], inf["ncols"], inf["nnzs"], inf["field"], inf["symmetric"]) == (n, n, len(v), 0, 1)
    A = ab.SymCsrMatrix.read_mtx(path, binary=binary)
    W = ab.SymCsrMatrix.init_real_double(n, r, c, v)
    _same(A, W)


@pytest.mark.parametrize("case", ["27pt-blocks", "rmat-metis"])
def test_read_part_matches_partition(case, ab, tmp_path):
    """acgb200_mtx_read_part(file, rowparts, p) == acgsymcsrmatrix_partition(whole matrix)[p],
    array for array, for every part."""
    if case == "27pt-blocks":
        n, r, c, v = mg.stencil3d_27pt(10, 9, 8)
        nparts = 6
        rowparts = abdist.block_partition(10, 9, 8, 2, 3, 1)
    else:
        n, r, c, v = mg.rmat_spd(3000, 20000, seed=9)
        nparts = 5
        rowparts, _ = ab.SymCsrMatrix.init_real_double(n, r, c, v).partition_rows(nparts, seed=1)
    v = v * (1 + 0.001 * np.arange(len(v)))
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=True)
    want = ab.SymCsrMatrix.init_real_double(n, r, c, v).partition(nparts, rowparts)
    for p in range(nparts):
        got = ab.SymCsrMatrix.read_mtx_part(path, nparts, rowparts, p)
        _same(got, want[p])
        got.free()


def test_read_part_spans_several_chunks(ab, tmp_path):
    """More entries than one read chunk (2^22): the kept entries are concatenated in
    file order, independent of the thread count."""
    n, r, c, v = mg.laplace3d_7pt(112, 112, 96)            # 4.8 M upper entries
    assert len(v) > 2 ** 22
    path = str(tmp_path / "A.mtx")
    mtxio.write_symmetric(path, n, r, c, v, binary=True)
    rowparts = abdist.block_partition(112, 112, 96, 2, 2, 2)
    got = ab.SymCsrMatrix.read_mtx_part(path, 8, rowparts, 5)
    want = ab.SymCsrMatrix.stencil_part(7, 112, 112, 96, 2, 2, 2, 5)
    _same(got, want)


def test_rejects_what_the_solver_cannot_take(ab, tmp_path):
    p = tmp_path / "g.mtx"
    p.write_bytes(b"%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 1.0\n")
    with pytest.raises(ab.AcgError) as e:
        ab.SymCsrMatrix.read_mtx(str(p), binary=False)
    assert e.value.code == ab.api.ACG_ERR_NOT_SUPPORTED
    p.write_bytes(b"%%MatrixMarket matrix coordinate real symmetric\n% a comment\n2 2 2\n1 1 1.0\n3 1 1.0\n")
    with pytest.raises(ab.AcgError) as e:
        ab.SymCsrMatrix.read_mtx(str(p), binary=False)
    assert e.value.code == ab.api.ACG_ERR_INDEX_OUT_OF_BOUNDS
    p.write_bytes(b"%%MatrixMarket matrix coordinate real symmetric\n2 2 3\n")      # truncated binary payload
    with pytest.raises(ab.AcgError):
        ab.SymCsrMatrix.read_mtx(str(p), binary=True)
    with pytest.raises(ab.AcgError):
        ab.SymCsrMatrix.read_mtx(str(tmp_path / "missing.mtx"))
