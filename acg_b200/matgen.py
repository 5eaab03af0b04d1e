"""Synthetic SPD test matrices for the CG path (host side, numpy).

All generators return the UPPER triangle (diagonal included) in coordinate
form, 0-based, row-sorted with ascending columns inside each row -- the input
convention of ``acgsymcsrmatrix_init_real_double`` (reference
acg/symcsrmatrix.c:66: "Only upper triangular entries of A should be
provided").  The configurations are the ones BASELINE.json / SURVEY.md §8(d)
name:

* ``poisson1d_5pt``  -- 4th-order 1-D Laplacian, diag 30, +-1 -> -16, +-2 -> +1
* ``laplace3d_7pt``  -- diag 6, six neighbours -1, Dirichlet truncation
* ``stencil3d_27pt`` -- diag 26, 26 neighbours -1 (HPCG style)
* ``rmat_spd``       -- R-MAT power-law graph Laplacian + I
* ``random_spd``     -- small dense-ish random SPD for known-answer tests

This module is data preparation, not the hot path; it never touches the GPU.
"""
from __future__ import annotations

import numpy as np


def _from_offsets(n, offsets, valid_fn, values):
    """Assemble row-sorted upper COO from per-offset masks.

    offsets: list of non-negative column offsets (ascending), values: value per
    offset, valid_fn(k) -> boolean mask of rows for which offset k exists.
    """
    nd = len(offsets)
    mask = np.empty((n, nd), dtype=bool)
    for k in range(nd):
        mask[:, k] = valid_fn(k)
    rows = np.broadcast_to(np.arange(n, dtype=np.int32)[:, None], (n, nd))[mask]
    cols = (np.arange(n, dtype=np.int64)[:, None] + np.asarray(offsets, dtype=np.int64)[None, :])[mask].astype(np.int32)
    vals = np.broadcast_to(np.asarray(values, dtype=np.float64)[None, :], (n, nd))[mask]
    return n, np.ascontiguousarray(rows), np.ascontiguousarray(cols), np.ascontiguousarray(vals)


def poisson1d_5pt(n: int):
    """1-D 5-point (4th order) Laplacian, SPD with Dirichlet truncation."""
    idx = np.arange(n)
    offsets = [0, 1, 2]
    values = [30.0, -16.0, 1.0]
    return _from_offsets(n, offsets, lambda k: idx + offsets[k] < n, values)


def poisson1d_3pt(n: int):
    """tridiag(-1, 2, -1): the closed-form known-answer case."""
    idx = np.arange(n)
    offsets = [0, 1]
    return _from_offsets(n, offsets, lambda k: idx + offsets[k] < n, [2.0, -1.0])


def _grid(nx, ny, nz):
    n = nx * ny * nz
    i = np.arange(n, dtype=np.int64)
    x = i % nx
    y = (i // nx) % ny
    z = i // (nx * ny)
    return n, x, y, z


def laplace3d_7pt(nx: int, ny: int | None = None, nz: int | None = None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n, x, y, z = _grid(nx, ny, nz)
    offsets = [0, 1, nx, nx * ny]
    masks = [np.ones(n, bool), x + 1 < nx, y + 1 < ny, z + 1 < nz]
    return _from_offsets(n, offsets, lambda k: masks[k], [6.0, -1.0, -1.0, -1.0])


def stencil3d_27pt(nx: int, ny: int | None = None, nz: int | None = None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n, x, y, z = _grid(nx, ny, nz)
    offs, masks = [], []
    for dz in (0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                o = dx + nx * dy + nx * ny * dz
                if o < 0 or (dz == 0 and dy == 0 and dx < 0):
                    continue
                if dz == 0 and dy < 0:
                    continue
                m = np.ones(n, bool)
                if dx < 0: m &= x >= 1
                if dx > 0: m &= x + 1 < nx
                if dy < 0: m &= y >= 1
                if dy > 0: m &= y + 1 < ny
                if dz > 0: m &= z + 1 < nz
                offs.append(o)
                masks.append(m)
    order = np.argsort(offs)
    offs = [offs[k] for k in order]
    masks = [masks[k] for k in order]
    vals = [26.0 if o == 0 else -1.0 for o in offs]
    return _from_offsets(n, offs, lambda k: masks[k], vals)


def rmat_spd(scale_n: int, nedges: int, seed: int = 42, abcd=(0.57, 0.19, 0.19, 0.05)):
    """R-MAT style power-law graph on ``scale_n`` vertices (rounded up to a
    power of two for edge generation, then folded with a modulo), symmetrised,
    deduplicated; A = D + I - Adj, strictly diagonally dominant => SPD."""
    rng = np.random.default_rng(seed)
    levels = int(np.ceil(np.log2(max(scale_n, 2))))
    a, b, c, _ = abcd
    src = np.zeros(nedges, dtype=np.int64)
    dst = np.zeros(nedges, dtype=np.int64)
    for _lvl in range(levels):
        u = rng.random(nedges)
        right = (u >= a) & (u < a + b) | (u >= a + b + c)
        down = u >= a + b
        src = (src << 1) | down
        dst = (dst << 1) | right
    src %= scale_n
    dst %= scale_n
    lo = np.minimum(src, dst)
    hi = np.maximum(src, dst)
    keep = lo != hi
    key = np.unique(lo[keep] * np.int64(scale_n) + hi[keep])
    lo = (key // scale_n).astype(np.int32)
    hi = (key % scale_n).astype(np.int32)
    deg = np.bincount(lo, minlength=scale_n) + np.bincount(hi, minlength=scale_n)
    rows = np.concatenate([np.arange(scale_n, dtype=np.int32), lo])
    cols = np.concatenate([np.arange(scale_n, dtype=np.int32), hi])
    vals = np.concatenate([deg.astype(np.float64) + 1.0, -np.ones(lo.size)])
    order = np.lexsort((cols, rows))
    return scale_n, rows[order], cols[order], vals[order]


def random_spd(n: int, density: float = 0.1, seed: int = 0):
    """Random symmetric, strictly diagonally dominant matrix (upper COO)."""
    rng = np.random.default_rng(seed)
    iu, ju = np.triu_indices(n, k=1)
    pick = rng.random(iu.size) < density
    iu, ju = iu[pick].astype(np.int32), ju[pick].astype(np.int32)
    off = rng.uniform(-1.0, 1.0, iu.size)
    rowsum = np.bincount(iu, weights=np.abs(off), minlength=n) + np.bincount(ju, weights=np.abs(off), minlength=n)
    diag = rowsum + rng.uniform(0.5, 1.5, n)
    rows = np.concatenate([np.arange(n, dtype=np.int32), iu])
    cols = np.concatenate([np.arange(n, dtype=np.int32), ju])
    vals = np.concatenate([diag, off])
    order = np.lexsort((cols, rows))
    return n, rows[order], cols[order], vals[order]


def upper_to_dense(n, rows, cols, vals):
    """Dense symmetric matrix from the upper COO (for tiny known-answer tests)."""
    A = np.zeros((n, n))
    A[rows, cols] = vals
    A[cols, rows] = vals
    return A


def upper_to_full_csr(n, rows, cols, vals):
    """Full-storage CSR (both triangles) from the upper COO, numpy only.

    Column order inside each row is ascending, which for row-sorted /
    column-ascending input equals the order the reference's
    acgsymcsrmatrix_dsymv_init produces (acg/symcsrmatrix.c:792-812).
    Returns (rowptr int64 [n+1], colidx int32, values float64).
    """
    off = rows != cols
    r = np.concatenate([rows, cols[off]]).astype(np.int64)
    c = np.concatenate([cols, rows[off]]).astype(np.int64)
    v = np.concatenate([vals, vals[off]])
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=rowptr[1:])
    return rowptr, c.astype(np.int32), v
