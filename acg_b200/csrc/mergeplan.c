/*
 * mergeplan.c -- host plan of the merge-path tiles of the SpMV (no CUDA calls).
 *
 * Power-law matrices (BASELINE config 5) defeat row-aligned tiles: a tile of 32-64 rows with one
 * 800-entry row keeps one lane group busy while the rest of the CTA idles, and the ncu capture of
 * round 2 shows exactly that on R-MAT 20 M (profiles/r02/b_ncu_rmat_tiles.json: 16 of 32 lanes
 * active, 36 % of the warps, DRAM at 16 % of peak although the traffic is the compulsory minimum).
 * The reference balances such rows with a merge-path SpMV (acg/cg-kernels-cuda.cu:312-441,
 * csrgemv_merge_startrows / csrgemv_merge); this is the same idea on top of the TMA ring:
 *
 * The row-end offsets rowptr[1..hi] and the nonzero indices 0..nnz-1 are merged ("consume nonzero
 * k while k < rowptr[r+1], else end row r"); the merged sequence of hi + nnz items is cut into
 * tiles of `items` consecutive items.  A tile therefore holds at most `items` nonzeros AND at most
 * `items` row ends, whatever the row lengths, and both slices are contiguous: one bulk copy each.
 * Tile t starts at the path point (r, k) on diagonal d = t * items: r is the largest row with
 * rowptr[r] + r <= d, k = d - r (then rowptr[r] <= k <= rowptr[r+1]).
 *
 * A row whose nonzeros lie in one tile is finished by that tile.  A row cut by tile boundaries is a
 * "split row": every tile it runs through writes the partial sum of its piece (slot 1: the piece at
 * the tile's end, which continues; slot 0: the piece at the tile's start, which ends there) and
 * spmv_merge_fix_kernel adds the pieces in tile order -- deterministic, no atomics on y.  Rows of
 * any length go through the same path (no separate long-row kernels below `hi`).
 */
#include "acgb200/error.h"
#include "internal.h"

#include <stdlib.h>
#include <string.h>

void acgb200_mergeplan_free(struct acgb200_mergeplan *mp)
{
    free(mp->tiles); free(mp->split);
    memset(mp, 0, sizeof(*mp));
}

/* path point on diagonal d: largest r in [0,hi] with rowptr[r] + r <= d */
static int path_row(const int64_t *rowptr, int hi, int64_t d)
{
    int lo = 0, up = hi;
    while (lo < up) {
        const int m = lo + (up - lo + 1) / 2;
        if (rowptr[m] + m <= d) lo = m; else up = m - 1;
    }
    return lo;
}

int acgb200_merge_plan(int hi, const int64_t *rowptr, int items, struct acgb200_mergeplan *out)
{
    memset(out, 0, sizeof(*out));
    if (hi <= 0 || items < 64) return ACG_SUCCESS;
    const int64_t nnz = rowptr[hi] - rowptr[0];
    if (rowptr[0] != 0 || rowptr[hi] > INT32_MAX) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    const int64_t total = (int64_t) hi + nnz;
    const int64_t nt64 = (total + items - 1) / items;
    if (nt64 > INT32_MAX / 2) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    const int nt = (int) nt64;
    out->tiles = malloc((size_t) (nt > 0 ? nt : 1) * sizeof(*out->tiles));
    out->split = malloc((size_t) (nt > 0 ? nt : 1) * sizeof(*out->split));
    if (!out->tiles || !out->split) { acgb200_mergeplan_free(out); return ACG_ERR_ERRNO; }
    int ns = 0;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; t++) {
        const int64_t d0 = (int64_t) t * items, d1 = d0 + items < total ? d0 + items : total;
        const int r0 = path_row(rowptr, hi, d0), r1 = path_row(rowptr, hi, d1);
        out->tiles[t].r0 = r0; out->tiles[t].nre = r1 - r0;
        out->tiles[t].k0 = (int) (d0 - r0); out->tiles[t].nnz = (int) ((d1 - r1) - (d0 - r0));
    }
    /* split rows: the row that ends in tile t but started before it */
    for (int t = 0; t < nt; t++) {
        const struct acgb200_mtile *tl = &out->tiles[t];
        if (tl->nre <= 0 || rowptr[tl->r0] >= tl->k0) continue;
        const int64_t dstart = rowptr[tl->r0] + tl->r0;         /* diagonal of the row's first nonzero */
        out->split[ns].row = tl->r0; out->split[ns].ta = (int) (dstart / items); out->split[ns].tb = t; out->split[ns].pad = 0;
        ns++;
    }
    out->ntiles = nt; out->nsplit = ns; out->items = items; out->rows = hi; out->nnz = nnz;
    return ACG_SUCCESS;
}
