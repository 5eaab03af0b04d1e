"""Host plan of the merge-path SpMV tiles (acg_b200/csrc/mergeplan.c), CPU only: the merged sequence of row
ends and nonzeros is cut into equal tiles; every nonzero and every row end belongs to exactly one tile; a
row cut by tile boundaries is listed as a split row with the tiles whose partial sums make it up; and an
emulation of spmv_merge_kernel + spmv_merge_fix_kernel from exactly these arrays gives y = A x.  (The
kernels themselves: tests/test_gpu_parity.py::test_merge_path_tiles_match_oracle on the B200; the solver
around them: tests/test_hostsim.py::test_power_law_rows on the device stand-in.)"""
import numpy as np
import pytest

from acg_b200 import matgen as mg


def _full(ab, gen):
    n, r, c, v = gen()
    A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
    return n, A.frowptr.copy(), A.fcolidx.copy(), A.fa.copy()


CASES = [("rmat-hubs", lambda: mg.rmat_spd(30000, 600000, seed=8)), ("27pt", lambda: mg.stencil3d_27pt(14)),
         ("1d3pt", lambda: mg.poisson1d_3pt(5000)), ("dense-rows", lambda: mg.random_spd(400, 0.5, 2))]


@pytest.mark.parametrize("items", [64, 256, 1024, 4096])
@pytest.mark.parametrize("name,gen", CASES, ids=[c[0] for c in CASES])
def test_tiles_partition_the_merged_sequence_and_reproduce_the_product(name, gen, items, ab):
    n, rp, col, val = _full(ab, gen)
    mp = ab.merge_plan_host(rp, items)
    tiles, split = mp["tiles"], mp["split"]
    nnz = int(rp[n])
    assert len(tiles) == (n + nnz + items - 1) // items
    # consecutive tiles continue each other; each holds at most `items` merged items
    r, k = 0, 0
    for r0, nre, k0, tn in tiles:
        assert (r0, k0) == (r, k) and 0 <= nre and 0 <= tn and nre + tn <= items
        assert rp[r0] <= k0 <= rp[r0 + 1] if r0 < n else k0 == nnz       # a valid point of the merge path
        r, k = r0 + nre, k0 + tn
    assert (r, k) == (n, nnz)
    # emulation: products per tile, units (row ends + the piece behind the last one), partial sums, fix-up
    x = np.random.default_rng(4).standard_normal(n)
    y = np.full(n, np.nan)
    part = np.zeros((len(tiles), 2))
    done = np.zeros(n, int)
    for t, (r0, nre, k0, tn) in enumerate(tiles):
        prod = val[k0:k0 + tn] * x[col[k0:k0 + tn]]
        kend = k0 + tn
        for j in range(nre + 1):
            a0 = max(rp[r0 + j], k0) - k0
            b0 = (rp[r0 + j + 1] if j < nre else kend) - k0
            s = prod[a0:b0].sum() if b0 > a0 else 0.0
            if j == nre:
                part[t, 1] = s                           # the row continues in the next tile
            elif j == 0 and rp[r0] < k0:
                part[t, 0] = s                           # the row began in an earlier tile and ends here
            else:
                y[r0 + j] = s; done[r0 + j] += 1
    for row, ta, tb in split:
        assert rp[row] < tiles[tb][2] and tiles[tb][0] == row and tiles[tb][1] > 0      # ends in tb, began before it
        assert (rp[row] + row) // items == ta < tb
        y[row] = part[ta:tb, 1].sum() + part[tb, 0]; done[row] += 1
    assert (done == 1).all()
    want = np.array([val[rp[i]:rp[i + 1]] @ x[col[rp[i]:rp[i + 1]]] for i in range(n)])
    scale = np.array([np.abs(val[rp[i]:rp[i + 1]]) @ np.abs(x[col[rp[i]:rp[i + 1]]]) for i in range(n)])
    assert np.all(np.abs(y - want) <= 1e-13 * scale + 1e-300)
    if name == "rmat-hubs" or items == 64:
        assert len(split) > 0


def test_rows_below_a_limit_only(ab):
    """Between GPUs the plan covers the interior rows only (hi = borderrowoffset)."""
    n, rp, col, val = _full(ab, CASES[0][1])
    hi = 20000
    mp = ab.merge_plan_host(rp, 512, hi=hi)
    r0, nre, k0, tn = mp["tiles"][-1]
    assert r0 + nre == hi and k0 + tn == rp[hi]
