"""Host plan of the SpMV's pattern slices (acg_b200/csrc/slices.c), CPU only.

The kernel that consumes the plan (spmv_slices_kernel) is tested on the B200
(tests/test_gpu_parity.py); the solver's host logic around it on the device stand-in
(tests/test_hostsim.py runs with the option on, its stand-in cross-checks every covered
row through the slice-major values against the CSR arrays)."""
import numpy as np
import pytest

from acg_b200 import matgen as mg


def _full(ab, gen):
    n, r, c, v = gen()
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    return n, A.frowptr.copy(), A.fcolidx.copy(), A.fa.copy()


@pytest.mark.parametrize("name,gen", [("27pt-12", lambda: mg.stencil3d_27pt(12)),
                                      ("27pt-aniso", lambda: mg.stencil3d_27pt(9, 40, 17)),
                                      ("7pt", lambda: mg.laplace3d_7pt(31, 17, 23)),
                                      ("1d5pt", lambda: mg.poisson1d_5pt(5000))])
def test_slices_cover_pattern_rows_exactly(name, gen, ab):
    n, rp, col, _ = _full(ab, gen)
    sp = ab.slices_host(rp, col)
    pat = ab.patterns_host(rp, col)
    assert sp["nslices"] > 0 and sp["rows"] == 32 * sp["nslices"]
    lens = np.diff(rp)
    lpad, table = sp["lpad"], sp["spatoff"]
    blocks = 0
    nnz = 0
    for row0, nrows, L, vblk in sp["slices"]:
        assert nrows == 32 and row0 % 32 == 0 and sp["covered"][row0 // 32]
        assert vblk == blocks
        assert L == lens[row0:row0 + 32].max() and L <= lpad
        blocks += L
        for r in range(row0, row0 + 32):
            pid = int(sp["patid"][r])
            assert pid != 0xFFFF and pid == int(pat["patid"][r])                  # whole stencils: no exception rows
            offs = table[pid * lpad:(pid + 1) * lpad]
            assert np.array_equal(r + offs[:lens[r]], col[rp[r]:rp[r + 1]])      # columns rebuilt from the table
            assert not offs[lens[r]:].any()                                       # padded slots gather x[row]
            nnz += lens[r]
    assert blocks == sp["blocks"] and nnz == sp["nnz"] and sp["nexc"] == 0
    # slices that are not covered: the ragged end (fewer than 32 rows) or too much padding
    assert sp["covered"].sum() == sp["nslices"]
    assert 8 * 32 * sp["blocks"] <= 11 * sp["nnz"]


def test_no_slices_for_unstructured_rows(ab):
    n, rp, col, _ = _full(ab, lambda: mg.random_spd(600, 0.05, 3))
    sp = ab.slices_host(rp, col)
    assert sp["nslices"] == 0 and not sp["covered"].any()


def test_exception_rows_inside_a_partitions_block(ab):
    """One block of a 2x2x2 partition: every grid line of the interior has a row next to the border shell whose
    offsets to the (separately numbered) border rows are unique -- not in the dictionary.  Such rows stay inside
    their slices as exception rows (columns from the index array): coverage stays near 100 % of the interior
    instead of losing every slice that holds one."""
    from acg_b200 import dist as abdist
    A = abdist.local_stencil_part(27, 48, 48, 48, 0, 8)
    no, bo = A.c.nownedrows, A.c.borderrowoffset
    rp = A.frowptr[:no + 1].copy(); col = A.fcolidx[:rp[no]].copy()
    sp = ab.slices_host(rp, col, cover_hi=bo)
    assert sp["nexc"] > 0 and sp["rows"] >= 0.95 * (bo - bo % 32)
    lens = np.diff(rp)
    nexc = excnnz = 0
    for row0, _, L, _ in sp["slices"]:
        ids = sp["patid"][row0:row0 + 32]
        assert (ids == 0xFFFF).sum() <= 8
        for r in np.nonzero(ids == 0xFFFF)[0] + row0:
            nexc += 1; excnnz += lens[r]
            assert lens[r] <= L
        for r in np.nonzero(ids != 0xFFFF)[0] + row0:
            offs = sp["spatoff"][int(sp["patid"][r]) * sp["lpad"]:][:sp["lpad"]]
            assert np.array_equal(r + offs[:lens[r]], col[rp[r]:rp[r + 1]])
    assert nexc == sp["nexc"] and excnnz == sp["excnnz"]


def test_cover_limit_excludes_border_rows(ab):
    """Between GPUs only rows below borderrowoffset may go to slices (the tile kernel adds the
    border x ghost block to the others)."""
    n, rp, col, _ = _full(ab, lambda: mg.stencil3d_27pt(16))
    hi = 2000
    sp = ab.slices_host(rp, col, cover_hi=hi)
    assert sp["nslices"] == hi // 32
    assert all(row0 + 32 <= hi for row0, *_ in sp["slices"])
    assert not sp["covered"][hi // 32:].any()
