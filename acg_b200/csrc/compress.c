/*
 * compress.c -- row-pattern dictionary (host side) of the SpMV's pattern slices.
 *
 * Matrices from stencils and other structured discretisations repeat a small
 * number of row "patterns" -- the sequence of offsets col - row of a row's
 * nonzeros: 27 of them describe every row of the 27-point stencil on a box.
 * A row whose pattern is in the dictionary needs no column indices at all: a
 * 2-byte pattern id per row replaces 4 bytes per nonzero (and gives the row's
 * length, so its row pointer goes too).  slices.c turns runs of 32 such rows
 * into slice-major storage for spmv_slices_kernel; rows with patterns outside
 * the dictionary (irregular matrices, rows next to a partition boundary after
 * the [interior|border] reordering) keep the CSR tiles.  Nothing here is
 * specific to stencils: the dictionary is found by hashing the rows of whatever
 * matrix is given, and an unstructured matrix simply ends up with no slices.
 *
 * History: round 1 used the dictionary for "index-free tiles" (the CSR tile
 * kernel rebuilding columns from the pattern in shared memory).  Measured on
 * the B200 in round 2 that kernel lost to the plain tiles (0.669 vs 0.631 ms at
 * C3: fewer bytes, but the tile kernel is instruction-bound at that point) and
 * was removed; the slice kernel is its successor.
 */
#include "acgb200/error.h"
#include "internal.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_num_threads(void) { return 1; }
#endif

struct slot { uint64_t hash; int row; int len; int64_t count; int id; };

static uint64_t row_hash(const int *col, int64_t kb, int64_t ke, int row)
{
    uint64_t h = 1469598103934665603ull ^ (uint64_t) (ke - kb);
    for (int64_t k = kb; k < ke; k++) {
        h ^= (uint64_t) (uint32_t) (col[k] - row);
        h *= 1099511628211ull;
    }
    return h ? h : 1;
}

static int same_pattern(const int64_t *rp, const int *col, int r1, int r2)
{
    const int64_t n1 = rp[r1 + 1] - rp[r1];
    if (n1 != rp[r2 + 1] - rp[r2]) return 0;
    for (int64_t j = 0; j < n1; j++)
        if (col[rp[r1] + j] - r1 != col[rp[r2] + j] - r2) return 0;
    return 1;
}

static int by_count_desc(const void *a, const void *b)
{
    const struct slot *x = a, *y = b;
    if (x->count != y->count) return x->count > y->count ? -1 : 1;
    return (x->row > y->row) - (x->row < y->row);       /* deterministic order */
}

void acgb200_patterns_free(struct acgb200_patterns *p)
{
    free(p->patptr); free(p->patoff); free(p->patid);
    memset(p, 0, sizeof(*p));
}

/*
 * Build the dictionary for rows [0,nrows) of a CSR matrix with 0-based
 * columns.  At most 65535 patterns totalling at most max_entries offsets are
 * kept, most frequent first.  patid[r] == ACGB200_NOPATTERN marks a row whose
 * pattern was not kept.
 */
/* count one row's pattern in an open-addressing table of `cap` slots holding at most cap/2 patterns;
 * returns 0 if the table is full (the pattern is then not tracked) */
static int table_count(struct slot *tab, size_t cap, size_t *used, const int64_t *rowptr, const int *colidx, int r, int64_t add, int rep)
{
    const uint64_t h = row_hash(colidx, rowptr[r], rowptr[r + 1], r);
    size_t i = (size_t) (h & (cap - 1));
    for (;;) {
        if (tab[i].hash == 0) {
            if (*used >= cap / 2) return 0;
            tab[i].hash = h; tab[i].row = rep; tab[i].len = (int) (rowptr[r + 1] - rowptr[r]);
            tab[i].count = add; tab[i].id = -1; (*used)++;
            return 1;
        }
        if (tab[i].hash == h && same_pattern(rowptr, colidx, tab[i].row, r)) {
            tab[i].count += add;
            if (rep < tab[i].row) tab[i].row = rep;      /* representative: the pattern's first row, whatever the thread count */
            return 1;
        }
        i = (i + 1) & (cap - 1);
    }
}

int acgb200_patterns_build(int nrows, const int64_t *rowptr, const int *colidx, int max_entries,
                           struct acgb200_patterns *out)
{
    memset(out, 0, sizeof(*out));
    const size_t cap = 1u << 17;          /* open addressing; distinct patterns tracked <= cap/2 */
    struct slot *tab = calloc(cap, sizeof(*tab));
    out->patid = malloc((size_t) (nrows > 0 ? nrows : 1) * sizeof(*out->patid));
    if (!tab || !out->patid) { free(tab); acgb200_patterns_free(out); return ACG_ERR_ERRNO; }
    size_t used = 0;
    /* pass 1: count patterns -- every thread counts a contiguous block of rows in a table of its own
     * (2^15 slots), the tables are merged in thread order */
    int failed = 0;
#pragma omp parallel
    {
        const size_t tcap = 1u << 15;
        struct slot *mine = calloc(tcap, sizeof(*mine));
        size_t mused = 0;
        if (!mine) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp for schedule(static)
        for (int r = 0; r < nrows; r++)
            if (mine) table_count(mine, tcap, &mused, rowptr, colidx, r, 1, r);
#pragma omp for ordered schedule(static, 1)
        for (int t = 0; t < omp_get_num_threads(); t++) {
#pragma omp ordered
            {
                if (mine)
                    for (size_t i = 0; i < tcap; i++)
                        if (mine[i].hash) table_count(tab, cap, &used, rowptr, colidx, mine[i].row, mine[i].count, mine[i].row);
            }
        }
        free(mine);
    }
    if (failed) { free(tab); acgb200_patterns_free(out); return ACG_ERR_ERRNO; }
    /* choose the most frequent patterns that fit */
    struct slot *sel = malloc((used ? used : 1) * sizeof(*sel));
    if (!sel) { free(tab); acgb200_patterns_free(out); return ACG_ERR_ERRNO; }
    size_t ns = 0;
    for (size_t i = 0; i < cap; i++) if (tab[i].hash) sel[ns++] = tab[i];
    qsort(sel, ns, sizeof(*sel), by_count_desc);
    int npat = 0, nent = 0;
    for (size_t i = 0; i < ns && npat < 65535; i++) {
        if (nent + sel[i].len > max_entries) continue;
        sel[i].id = npat++; nent += sel[i].len;
    }
    out->npat = npat; out->nentries = nent;
    out->patptr = malloc(((size_t) npat + 1) * sizeof(int));
    out->patoff = malloc((size_t) (nent > 0 ? nent : 1) * sizeof(int));
    if (!out->patptr || !out->patoff) { free(sel); free(tab); acgb200_patterns_free(out); return ACG_ERR_ERRNO; }
    int pos = 0;
    for (size_t i = 0; i < ns; i++) {
        if (sel[i].id < 0) continue;
        out->patptr[sel[i].id] = pos;
        const int r = sel[i].row;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) out->patoff[pos++] = colidx[k] - r;
        /* write the id back into the hash table entry of this pattern */
        size_t j = (size_t) (sel[i].hash & (cap - 1));
        while (!(tab[j].hash == sel[i].hash && tab[j].row == sel[i].row)) j = (j + 1) & (cap - 1);
        tab[j].id = sel[i].id;
    }
    out->patptr[npat] = pos;
    free(sel);
    /* pass 2: ids per row */
    int64_t matched = 0;
#pragma omp parallel for schedule(static) reduction(+ : matched)
    for (int r = 0; r < nrows; r++) {
        const uint64_t h = row_hash(colidx, rowptr[r], rowptr[r + 1], r);
        size_t i = (size_t) (h & (cap - 1));
        int id = -1;
        while (tab[i].hash) {
            if (tab[i].hash == h && same_pattern(rowptr, colidx, tab[i].row, r)) { id = tab[i].id; break; }
            i = (i + 1) & (cap - 1);
        }
        out->patid[r] = id >= 0 ? (unsigned short) id : ACGB200_NOPATTERN;
        matched += id >= 0;
    }
    out->nrows_matched = matched;
    free(tab);
    return ACG_SUCCESS;
}
