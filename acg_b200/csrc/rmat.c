/*
 * rmat.c -- R-MAT style power-law SPD test matrix, generated in parallel.
 *
 * BASELINE.json config 5 ("power-law synthetic SPD CSR n=20M nnz~400M,
 * R-MAT-style"; SURVEY.md section 8d input 5): n vertices, `nedges` directed
 * R-MAT draws with quadrant probabilities (a,b,c,d), folded to n by a modulo,
 * self-loops dropped, symmetrised and deduplicated; A = D + I - Adj, i.e.
 * off-diagonal -1 and diagonal degree+1: strictly diagonally dominant, hence
 * SPD.  This exercises what stencils do not: rows of very different lengths
 * (the long-row path of the SpMV) and gathers without locality.
 *
 * Every edge comes from a counter-based generator (splitmix64 of seed, edge
 * number and recursion level), so the matrix does not depend on the number of
 * threads; edges are bucketed by their smaller end, each bucket is sorted and
 * deduplicated, and the packed upper triangle is handed to
 * acgsymcsrmatrix_init_rowwise_real_double.  20 M vertices / 200 M draws take
 * well under a minute on a multi-core host (the numpy generator of
 * acg_b200/matgen.py, used for the small test cases, needs ~8 minutes).
 */
#include "acgb200/error.h"
#include "acgb200/ext.h"
#include "acgb200/symcsrmatrix.h"
#include "hostmem.h"

#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static int cmp_idx(const void *a, const void *b)
{
    const acgidx_t x = *(const acgidx_t *) a, y = *(const acgidx_t *) b;
    return (x > y) - (x < y);
}

/* one R-MAT draw -> (lo, hi) with lo < hi, or lo == hi for a self-loop */
static inline void rmat_edge(uint64_t seed, int64_t e, int levels, int64_t n, const double *abcd,
                             acgidx_t *lo, acgidx_t *hi)
{
    const double a = abcd[0], ab = abcd[0] + abcd[1], abc = abcd[0] + abcd[1] + abcd[2];
    uint64_t src = 0, dst = 0;
    uint64_t state = splitmix64(seed ^ splitmix64((uint64_t) e));
    for (int l = 0; l < levels; l++) {
        state = splitmix64(state);
        const double u = (double) (state >> 11) * (1.0 / 9007199254740992.0);
        const int right = (u >= a && u < ab) || u >= abc;
        const int down = u >= ab;
        src = (src << 1) | (uint64_t) down;
        dst = (dst << 1) | (uint64_t) right;
    }
    const int64_t s = (int64_t) (src % (uint64_t) n), d = (int64_t) (dst % (uint64_t) n);
    *lo = (acgidx_t) (s < d ? s : d);
    *hi = (acgidx_t) (s < d ? d : s);
}

int acgb200_rmat_spd(int64_t n, int64_t nedges, uint64_t seed, const double *abcd, struct acgsymcsrmatrix *A)
{
    static const double default_abcd[4] = { 0.57, 0.19, 0.19, 0.05 };
    if (!abcd) abcd = default_abcd;
    if (n < 2 || n > ACGIDX_T_MAX || nedges < 0) return ACG_ERR_INVALID_VALUE;
    int levels = 1;
    while (((int64_t) 1 << levels) < n) levels++;

    int err = ACG_ERR_ERRNO;
    int64_t *bptr = acgb200_bigcalloc((size_t) n + 1, sizeof(*bptr));      /* bucket = smaller end */
    int64_t *cur = NULL, *rowptr = NULL;
    acgidx_t *bucket = NULL, *colidx = NULL;
    int *deg = NULL;
    double *vals = NULL;
    if (!bptr) goto done;

    /* pass 1: bucket sizes */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nedges; e++) {
        acgidx_t lo, hi;
        rmat_edge(seed, e, levels, n, abcd, &lo, &hi);
        if (lo != hi) __atomic_fetch_add(&bptr[lo + 1], 1, __ATOMIC_RELAXED);
    }
    for (int64_t i = 0; i < n; i++) bptr[i + 1] += bptr[i];
    const int64_t kept = bptr[n];
    bucket = acgb200_bigalloc((size_t) (kept > 0 ? kept : 1) * sizeof(*bucket));
    cur = acgb200_bigalloc((size_t) n * sizeof(*cur));
    if (!bucket || !cur) goto done;
    memcpy(cur, bptr, (size_t) n * sizeof(*cur));

    /* pass 2: the same draws again, larger ends into the buckets (any order) */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nedges; e++) {
        acgidx_t lo, hi;
        rmat_edge(seed, e, levels, n, abcd, &lo, &hi);
        if (lo != hi) bucket[__atomic_fetch_add(&cur[lo], 1, __ATOMIC_RELAXED)] = hi;
    }

    /* sort + deduplicate every bucket (this fixes the order); count the degrees */
    deg = calloc((size_t) n, sizeof(*deg));
    if (!deg) goto done;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; i++) {
        acgidx_t *b = bucket + bptr[i];
        const int64_t m = bptr[i + 1] - bptr[i];
        if (m > 1) qsort(b, (size_t) m, sizeof(*b), cmp_idx);
        int64_t u = 0;
        for (int64_t k = 0; k < m; k++) {
            if (k > 0 && b[k] == b[k - 1]) continue;
            b[u++] = b[k];
            __atomic_fetch_add(&deg[b[k]], 1, __ATOMIC_RELAXED);
        }
        cur[i] = u;                                  /* unique neighbours with a larger index */
        __atomic_fetch_add(&deg[i], (int) u, __ATOMIC_RELAXED);
    }

    /* packed upper triangle: diagonal first, then the larger neighbours ascending */
    rowptr = acgb200_bigalloc(((size_t) n + 1) * sizeof(*rowptr));
    if (!rowptr) goto done;
    rowptr[0] = 0;
    for (int64_t i = 0; i < n; i++) rowptr[i + 1] = rowptr[i] + 1 + cur[i];
    const int64_t nnz = rowptr[n];
    colidx = acgb200_bigalloc((size_t) nnz * sizeof(*colidx));
    vals = acgb200_bigalloc((size_t) nnz * sizeof(*vals));
    if (!colidx || !vals) goto done;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int64_t k = rowptr[i];
        colidx[k] = (acgidx_t) i; vals[k] = (double) deg[i] + 1.0; k++;
        const acgidx_t *b = bucket + bptr[i];
        for (int64_t j = 0; j < cur[i]; j++, k++) { colidx[k] = b[j]; vals[k] = -1.0; }
    }
    err = acgsymcsrmatrix_init_rowwise_real_double(A, (acgidx_t) n, 0, rowptr, colidx, vals);
done:
    free(bptr); free(cur); free(bucket); free(deg); free(rowptr); free(colidx); free(vals);
    return err;
}
