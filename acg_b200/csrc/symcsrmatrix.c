/*
 * symcsrmatrix.c -- host-side symmetric CSR matrices and their row partition.
 *
 * Own implementation of the acgsymcsrmatrix_* entry points declared in
 * include/acgb200/symcsrmatrix.h.  It produces the data layout that the
 * device solver (cgcuda.c) consumes and that the reference documents in
 * acg/symcsrmatrix.h:62-292 / acg/graph.h:54-329:
 *
 *   packed storage   upper triangle, CSR by row            (symcsrmatrix.c:66)
 *   full storage     local block (both triangles) + border x ghost block
 *                                                          (symcsrmatrix.c:760)
 *   partition        local order [interior | border | ghost], ghosts grouped
 *                    by owning part and ascending in global index, per
 *                    neighbour border/ghost lists           (graph.c:813-1446)
 *
 * The algorithms are not the reference's (which goes through a general graph
 * partitioning layer with radix sorts and edge-orientation flags); they are
 * direct constructions of the same invariants, checked against the reference
 * build in tests/test_partition_pin.py.
 */
#include "acgb200/error.h"
#include "acgb200/halo.h"
#include "acgb200/symcsrmatrix.h"
#include "hostmem.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* graph bookkeeping                                                         */
/* ------------------------------------------------------------------------ */

static void graph_free(struct acggraph *g)
{
    if (!g) return;
    free(g->nodetags); free(g->parentnodeidx);
    free(g->edgetags); free(g->parentedgeidx);
    free(g->nodenedges); free(g->srcnodeptr); free(g->srcnodeidx); free(g->dstnodeidx);
    free(g->nbordernodeinneredges); free(g->nbordernodeinterfaceedges);
    for (int i = 0; i < g->nneighbours; i++) {
        free(g->neighbours[i].bordernodes);
        free(g->neighbours[i].ghostnodes);
    }
    free(g->neighbours);
    memset(g, 0, sizeof(*g));
}

/* point the matrix's read-only views at the arrays its graph owns
 * (the aliasing of acg/symcsrmatrix.c:104-123) */
static void matrix_view_graph(struct acgsymcsrmatrix *A)
{
    const struct acggraph *g = A->graph;
    A->nrows = g->nnodes; A->nprows = g->npnodes; A->nzrows = g->parentnodeidx;
    A->nnzs = g->nedges; A->npnzs = g->npedges;
    A->rowidxbase = g->nodeidxbase;
    A->rownnzs = g->nodenedges; A->rowptr = g->srcnodeptr;
    A->rowidx = g->srcnodeidx; A->colidx = g->dstnodeidx;
    A->nownedrows = g->nownednodes; A->ninnerrows = g->ninnernodes;
    A->nborderrows = g->nbordernodes; A->borderrowoffset = g->bordernodeoffset;
    A->nghostrows = g->nghostnodes; A->ghostrowoffset = g->ghostnodeoffset;
    A->ninnernzs = g->ninneredges; A->ninterfacenzs = g->ninterfaceedges;
    A->nborderrowinnernzs = g->nbordernodeinneredges;
    A->nborderrowinterfacenzs = g->nbordernodeinterfaceedges;
    A->fnpnzs = A->onpnzs = 0;
    A->frowptr = A->orowptr = NULL;
    A->fcolidx = A->ocolidx = NULL;
    A->fa = A->oa = NULL;
}

/* a whole (unpartitioned) graph from CSR arrays it takes ownership of */
static struct acggraph *graph_adopt_csr(acgidx_t n, int idxbase, int64_t *rowptr, acgidx_t *colidx)
{
    struct acggraph *g = calloc(1, sizeof(*g));
    if (!g) return NULL;
    const int64_t nnz = rowptr[n];
    g->nparts = 1; g->parttag = 1; g->nprocs = 1; g->npparts = 1;
    g->nnodes = g->npnodes = n;
    g->nedges = g->npedges = nnz;
    g->nodeidxbase = idxbase;
    g->srcnodeptr = rowptr;
    g->dstnodeidx = colidx;
    g->nodenedges = malloc((size_t) (n > 0 ? n : 1) * sizeof(*g->nodenedges));
    g->srcnodeidx = malloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*g->srcnodeidx));
    if (!g->nodenedges || !g->srcnodeidx) { graph_free(g); free(g); return NULL; }
    #pragma omp parallel for
    for (acgidx_t i = 0; i < n; i++) {
        g->nodenedges[i] = rowptr[i + 1] - rowptr[i];
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) g->srcnodeidx[k] = i + idxbase;
    }
    g->nownednodes = g->ninnernodes = n;
    g->bordernodeoffset = g->ghostnodeoffset = n;
    g->ninneredges = nnz;
    return g;
}

/* ------------------------------------------------------------------------ */
/* construction                                                              */
/* ------------------------------------------------------------------------ */

int acgsymcsrmatrix_init_rowwise_real_double(
    struct acgsymcsrmatrix *A, acgidx_t N, int idxbase,
    const int64_t *rowptr, const acgidx_t *colidx, const double *a)
{
    const int64_t nnz = rowptr[N];
    int64_t *rp = malloc(((size_t) N + 1) * sizeof(*rp));
    acgidx_t *ci = malloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*ci));
    double *va = malloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*va));
    if (!rp || !ci || !va) { free(rp); free(ci); free(va); return ACG_ERR_ERRNO; }
    memcpy(rp, rowptr, ((size_t) N + 1) * sizeof(*rp));
    memcpy(ci, colidx, (size_t) nnz * sizeof(*ci));
    memcpy(va, a, (size_t) nnz * sizeof(*va));
    struct acggraph *g = graph_adopt_csr(N, idxbase, rp, ci);
    if (!g) { free(va); return ACG_ERR_ERRNO; }
    memset(A, 0, sizeof(*A));
    A->graph = g;
    matrix_view_graph(A);
    A->a = va;
    return ACG_SUCCESS;
}

/* Upper-triangle COO -> packed CSR by a stable counting sort on the row index,
 * i.e. entries of one row keep their input order, as in
 * acg/symcsrmatrix.c:84-96 / :133-141. */
int acgsymcsrmatrix_init_real_double(
    struct acgsymcsrmatrix *A, acgidx_t N, int64_t nnzs, int idxbase,
    const acgidx_t *rowidx, const acgidx_t *colidx, const double *a)
{
    for (int64_t k = 0; k < nnzs; k++) {
        if (rowidx[k] < idxbase || rowidx[k] >= N + idxbase ||
            colidx[k] < idxbase || colidx[k] >= N + idxbase)
            return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    }
    int64_t *rp = calloc((size_t) N + 1, sizeof(*rp));
    acgidx_t *ci = malloc((size_t) (nnzs > 0 ? nnzs : 1) * sizeof(*ci));
    double *va = malloc((size_t) (nnzs > 0 ? nnzs : 1) * sizeof(*va));
    int64_t *cur = malloc((size_t) (N > 0 ? N : 1) * sizeof(*cur));
    if (!rp || !ci || !va || !cur) { free(rp); free(ci); free(va); free(cur); return ACG_ERR_ERRNO; }
    for (int64_t k = 0; k < nnzs; k++) rp[rowidx[k] - idxbase + 1]++;
    for (acgidx_t i = 0; i < N; i++) rp[i + 1] += rp[i];
    memcpy(cur, rp, (size_t) N * sizeof(*cur));
    for (int64_t k = 0; k < nnzs; k++) {
        const int64_t l = cur[rowidx[k] - idxbase]++;
        ci[l] = colidx[k];
        va[l] = a[k];
    }
    free(cur);
    struct acggraph *g = graph_adopt_csr(N, idxbase, rp, ci);
    if (!g) { free(va); return ACG_ERR_ERRNO; }
    memset(A, 0, sizeof(*A));
    A->graph = g;
    matrix_view_graph(A);
    A->a = va;
    return ACG_SUCCESS;
}

void acgsymcsrmatrix_free(struct acgsymcsrmatrix *A)
{
    if (A->graph) { graph_free(A->graph); free(A->graph); }
    free(A->a);
    free(A->frowptr); free(A->fcolidx); free(A->fa);
    free(A->orowptr); free(A->ocolidx); free(A->oa);
    memset(A, 0, sizeof(*A));
}

int acgsymcsrmatrix_vector(const struct acgsymcsrmatrix *A, struct acgvector *x)
{
    /* acg/symcsrmatrix.c:634-645 */
    if (!A->nzrows) return acgvector_alloc(x, A->nrows);
    int err = acgvector_alloc_packed(x, A->nrows, A->nprows, 0, A->nzrows);
    if (err) return err;
    x->num_ghost_nonzeros = A->nghostrows;
    return ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* full storage                                                              */
/* ------------------------------------------------------------------------ */

/*
 * Packed upper triangle -> (a) local block with both triangles, restricted to
 * columns below ghostrowoffset, diagonal shifted by eps; (b) border x ghost
 * block, columns rebased by -borderrowoffset.  Entry order inside a full row
 * is the reference's (acg/symcsrmatrix.c:792-812): packed rows are visited in
 * increasing order and each entry (i,j) is appended to row i and, if i != j,
 * to row j -- so tests can compare the arrays verbatim.
 *
 * Full-storage expansion, threaded.  The result is the one the serial fill
 * produces (row r holds its entries ordered by the packed row they came from:
 * mirrored entries (i,r) at position i, its own packed entries at position r),
 * so the order of the products in every SpMV row -- and hence every bit of the
 * GPU result -- does not depend on the thread count:
 *   1. counts: own entries per row directly, mirrored ones with atomic adds;
 *      per packed row the [min,max] local-block column is recorded;
 *   2. serial prefix sum;
 *   3. fill: thread t owns a contiguous range of target rows (balanced by
 *      entries) and walks the packed rows in order, taking row i itself when i is
 *      in its range and the mirrored entries whose column falls in the range;
 *      rows whose column span misses the range are skipped on the two recorded
 *      bounds, so a banded matrix is read about once in total.
 */
int acgsymcsrmatrix_dsymv_init(struct acgsymcsrmatrix *A, double eps)
{
    const acgidx_t n = A->nprows, ghost0 = A->ghostrowoffset, border0 = A->borderrowoffset;
    const int base = A->rowidxbase;
    const int64_t *rp = A->rowptr;
    const acgidx_t *cj = A->colidx;
    const double *av = A->a;
    free(A->frowptr); free(A->fcolidx); free(A->fa);
    free(A->orowptr); free(A->ocolidx); free(A->oa);
    A->fcolidx = NULL; A->fa = NULL; A->orowptr = NULL; A->ocolidx = NULL; A->oa = NULL;
    A->frowptr = acgb200_bigcalloc((size_t) n + 1, sizeof(*A->frowptr));
    acgidx_t *span = acgb200_bigalloc((size_t) (n > 0 ? n : 1) * 2 * sizeof(*span));
    if (!A->frowptr || !span) { free(span); return ACG_ERR_ERRNO; }
    int64_t *cnt = A->frowptr + 1;
#pragma omp parallel for schedule(static)
    for (acgidx_t i = 0; i < n; i++) {
        int64_t own = 0;
        acgidx_t lo = ghost0, hi = -1;
        for (int64_t k = rp[i]; k < rp[i + 1]; k++) {
            const acgidx_t j = cj[k] - base;
            if (j >= ghost0) continue;
            own++;
            if (j < lo) lo = j;
            if (j > hi) hi = j;
            if (j != i) __atomic_fetch_add(&cnt[j], 1, __ATOMIC_RELAXED);
        }
        if (own) __atomic_fetch_add(&cnt[i], own, __ATOMIC_RELAXED);
        span[2 * (size_t) i] = lo; span[2 * (size_t) i + 1] = hi;
    }
    for (acgidx_t i = 0; i < n; i++) A->frowptr[i + 1] += A->frowptr[i];
    A->fnpnzs = A->frowptr[n];
    const size_t fcap = (size_t) (A->fnpnzs > 0 ? A->fnpnzs : 1);
    A->fcolidx = acgb200_bigalloc(fcap * sizeof(*A->fcolidx));
    A->fa = acgb200_bigalloc(fcap * sizeof(*A->fa));
    int64_t *cur = acgb200_bigalloc((size_t) (n > 0 ? n : 1) * sizeof(*cur));
    if (!A->fcolidx || !A->fa || !cur) { free(cur); free(span); return ACG_ERR_ERRNO; }
    memcpy(cur, A->frowptr, (size_t) n * sizeof(*cur));
    const int64_t *frp = A->frowptr;
    acgidx_t *fcj = A->fcolidx;
    double *fav = A->fa;
#pragma omp parallel
    {
        int nt = 1, t = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads(); t = omp_get_thread_num();
#endif
        /* target rows [ra, rb): equal shares of the full-storage entries */
        acgidx_t ra = 0, rb = n;
        if (nt > 1) {
            const int64_t tot = frp[n];
            const int64_t wa = tot / nt * t, wb = (t == nt - 1) ? tot : tot / nt * (t + 1);
            acgidx_t lo = 0, hi = n;
            while (lo < hi) { acgidx_t m = lo + (hi - lo) / 2; if (frp[m] < wa) lo = m + 1; else hi = m; }
            ra = lo; hi = n;
            while (lo < hi) { acgidx_t m = lo + (hi - lo) / 2; if (frp[m] < wb) lo = m + 1; else hi = m; }
            rb = (t == nt - 1) ? n : lo;
        }
        for (acgidx_t i = 0; i < n && ra < rb; i++) {
            const int mine = i >= ra && i < rb;
            if (!mine && (span[2 * (size_t) i + 1] < ra || span[2 * (size_t) i] >= rb)) continue;
            for (int64_t k = rp[i]; k < rp[i + 1]; k++) {
                const acgidx_t j = cj[k] - base;
                if (j >= ghost0) continue;
                if (mine) {
                    const int64_t l = cur[i]++;
                    fcj[l] = j + base;
                    fav[l] = av[k] + (i == j ? eps : 0.0);
                }
                if (j != i && j >= ra && j < rb) {
                    const int64_t l = cur[j]++;
                    fcj[l] = i + base;
                    fav[l] = av[k];
                }
            }
        }
    }
    free(cur); free(span);

    const acgidx_t no = A->nborderrows + A->nghostrows;
    A->orowptr = calloc((size_t) no + 1, sizeof(*A->orowptr));
    if (!A->orowptr) return ACG_ERR_ERRNO;
#pragma omp parallel for schedule(static)
    for (acgidx_t i = 0; i < no; i++) {
        int64_t c = 0;
        for (int64_t k = rp[border0 + i]; k < rp[border0 + i + 1]; k++)
            if (cj[k] - base >= ghost0) c++;
        A->orowptr[i + 1] = c;
    }
    for (acgidx_t i = 0; i < no; i++) A->orowptr[i + 1] += A->orowptr[i];
    A->onpnzs = A->orowptr[no];
    const size_t ocap = (size_t) (A->onpnzs > 0 ? A->onpnzs : 1);
    A->ocolidx = malloc(ocap * sizeof(*A->ocolidx));
    A->oa = malloc(ocap * sizeof(*A->oa));
    if (!A->ocolidx || !A->oa) return ACG_ERR_ERRNO;
#pragma omp parallel for schedule(static)
    for (acgidx_t i = 0; i < no; i++) {
        int64_t l = A->orowptr[i];
        for (int64_t k = rp[border0 + i]; k < rp[border0 + i + 1]; k++) {
            const acgidx_t j = cj[k] - base;
            if (j < ghost0) continue;
            A->ocolidx[l] = j + base - border0;
            A->oa[l] = av[k];
            l++;
        }
    }
    return ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* row partition                                                             */
/* ------------------------------------------------------------------------ */

struct pair { acgidx_t key; int part; };   /* (global node, other part) */

static int cmp_part_key(const void *a, const void *b)
{
    const struct pair *x = a, *y = b;
    if (x->part != y->part) return x->part < y->part ? -1 : 1;
    return (x->key > y->key) - (x->key < y->key);
}

static size_t uniq_pairs(struct pair *v, size_t n)
{
    if (n == 0) return 0;
    size_t m = 1;
    for (size_t i = 1; i < n; i++)
        if (v[i].key != v[m - 1].key || v[i].part != v[m - 1].part) v[m++] = v[i];
    return m;
}

/* position of (key) among the ghosts owned by `part` in a list sorted by (part,key) */
static acgidx_t find_pair(const struct pair *v, size_t lo, size_t hi, acgidx_t key)
{
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (v[mid].key < key) lo = mid + 1; else hi = mid;
    }
    return (acgidx_t) lo;
}

/*
 * Split A by the row->part map into nparts submatrices.  For part p:
 *   owned rows    = rows mapped to p; "border" if adjacent (in the symmetric
 *                   pattern) to a row of another part, else "interior"
 *   ghost rows    = rows of other parts adjacent to an owned row
 *   local order   = interior (ascending global), border (ascending global),
 *                   ghosts grouped by owner part (ascending), ascending global
 *   packed edges  = every edge with both ends owned (stored at the row of its
 *                   global row index) and every border-ghost edge, stored at
 *                   the border row with the ghost as column, so that
 *                   acgsymcsrmatrix_dsymv_init routes it to the off-diagonal
 *                   block
 *   neighbours    = one per adjacent part q (ascending): the border rows
 *                   adjacent to q and the ghosts owned by q, both ascending
 *                   in global index -- hence q's send list to p enumerates
 *                   exactly p's ghost segment for q, in order.
 */
int acgsymcsrmatrix_partition(
    const struct acgsymcsrmatrix *A, int nparts, const int *rowparts,
    struct acgsymcsrmatrix *sub, int verbose)
{
    (void) verbose;
    const acgidx_t n = A->nprows;
    const int base = A->rowidxbase;
    const int64_t *rp = A->rowptr;
    const acgidx_t *cj = A->colidx;
    int err = ACG_ERR_ERRNO;
    int64_t **cur = NULL;           /* scratch of pass 2, released on every path */
    size_t **gseg = NULL;
    if (A->nzrows || A->nghostrows) return ACG_ERR_NOT_SUPPORTED;   /* partition whole matrices only */
    for (acgidx_t i = 0; i < n; i++)
        if (rowparts[i] < 0 || rowparts[i] >= nparts) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    memset(sub, 0, (size_t) nparts * sizeof(*sub));      /* the failure path looks at every part */

    /* pass 1: cut edges -> (node, other part) pairs for both ends */
    size_t *ncut = calloc((size_t) nparts + 1, sizeof(*ncut));
    unsigned char *isborder = calloc((size_t) (n > 0 ? n : 1), 1);
    struct pair **ghosts = calloc((size_t) nparts, sizeof(*ghosts));
    struct pair **borders = calloc((size_t) nparts, sizeof(*borders));
    size_t *nghostp = calloc((size_t) nparts, sizeof(*nghostp));
    size_t *nborderp = calloc((size_t) nparts, sizeof(*nborderp));
    acgidx_t *localof = malloc((size_t) (n > 0 ? n : 1) * sizeof(*localof));
    acgidx_t *nowned = calloc((size_t) nparts, sizeof(*nowned));
    acgidx_t *ninner = calloc((size_t) nparts, sizeof(*ninner));
    int64_t *nedgesp = calloc((size_t) nparts, sizeof(*nedgesp));
    if (!ncut || !isborder || !ghosts || !borders || !nghostp || !nborderp || !localof || !nowned || !ninner || !nedgesp)
        goto fail;
    for (acgidx_t u = 0; u < n; u++) {
        const int pu = rowparts[u];
        for (int64_t k = rp[u]; k < rp[u + 1]; k++) {
            const acgidx_t v = cj[k] - base;
            const int pv = rowparts[v];
            if (pu == pv) { nedgesp[pu]++; continue; }
            ncut[pu]++; ncut[pv]++;
            nedgesp[pu]++; nedgesp[pv]++;
            isborder[u] = isborder[v] = 1;
        }
    }
    for (int p = 0; p < nparts; p++) {
        ghosts[p] = malloc((ncut[p] ? ncut[p] : 1) * sizeof(struct pair));
        borders[p] = malloc((ncut[p] ? ncut[p] : 1) * sizeof(struct pair));
        if (!ghosts[p] || !borders[p]) goto fail;
    }
    for (acgidx_t u = 0; u < n; u++) {
        const int pu = rowparts[u];
        for (int64_t k = rp[u]; k < rp[u + 1]; k++) {
            const acgidx_t v = cj[k] - base;
            const int pv = rowparts[v];
            if (pu == pv) continue;
            ghosts[pu][nghostp[pu]++] = (struct pair) { v, pv };
            borders[pu][nborderp[pu]++] = (struct pair) { u, pv };
            ghosts[pv][nghostp[pv]++] = (struct pair) { u, pu };
            borders[pv][nborderp[pv]++] = (struct pair) { v, pu };
        }
    }
    for (int p = 0; p < nparts; p++) {
        qsort(ghosts[p], nghostp[p], sizeof(struct pair), cmp_part_key);
        nghostp[p] = uniq_pairs(ghosts[p], nghostp[p]);
        qsort(borders[p], nborderp[p], sizeof(struct pair), cmp_part_key);
        nborderp[p] = uniq_pairs(borders[p], nborderp[p]);
    }

    /* local numbering of owned rows: interior first, then border, ascending global */
    for (acgidx_t u = 0; u < n; u++) { nowned[rowparts[u]]++; if (!isborder[u]) ninner[rowparts[u]]++; }
    {
        acgidx_t *nexti = calloc((size_t) nparts, sizeof(*nexti));
        acgidx_t *nextb = calloc((size_t) nparts, sizeof(*nextb));
        if (!nexti || !nextb) { free(nexti); free(nextb); goto fail; }
        for (acgidx_t u = 0; u < n; u++) {
            const int p = rowparts[u];
            localof[u] = isborder[u] ? ninner[p] + nextb[p]++ : nexti[p]++;
        }
        free(nexti); free(nextb);
    }

    /* pass 2: build every part */
    for (int p = 0; p < nparts; p++) {
        struct acgsymcsrmatrix *S = &sub[p];
        memset(S, 0, sizeof(*S));
        struct acggraph *g = calloc(1, sizeof(*g));
        if (!g) goto fail;
        S->graph = g;
        const acgidx_t no = nowned[p], ni = ninner[p], nb = no - ni, ng = (acgidx_t) nghostp[p];
        const acgidx_t npn = no + ng;
        g->nparts = nparts; g->parttag = p + 1; g->nprocs = nparts; g->npparts = 1;
        g->ownerrank = p; g->ownerpart = 0;
        g->nnodes = A->nrows; g->npnodes = npn;
        g->nedges = A->nnzs; g->npedges = nedgesp[p];
        g->nodeidxbase = 0;
        g->nownednodes = no; g->ninnernodes = ni; g->nbordernodes = nb; g->bordernodeoffset = ni;
        g->nghostnodes = ng; g->ghostnodeoffset = no;
        g->parentnodeidx = malloc((size_t) (npn > 0 ? npn : 1) * sizeof(*g->parentnodeidx));
        g->srcnodeptr = calloc((size_t) npn + 1, sizeof(*g->srcnodeptr));
        g->nodenedges = calloc((size_t) (npn > 0 ? npn : 1), sizeof(*g->nodenedges));
        g->srcnodeidx = malloc((size_t) (nedgesp[p] > 0 ? nedgesp[p] : 1) * sizeof(*g->srcnodeidx));
        g->dstnodeidx = malloc((size_t) (nedgesp[p] > 0 ? nedgesp[p] : 1) * sizeof(*g->dstnodeidx));
        g->nbordernodeinneredges = calloc((size_t) (nb > 0 ? nb : 1), sizeof(int64_t));
        g->nbordernodeinterfaceedges = calloc((size_t) (nb > 0 ? nb : 1), sizeof(int64_t));
        S->a = malloc((size_t) (nedgesp[p] > 0 ? nedgesp[p] : 1) * sizeof(*S->a));
        if (!g->parentnodeidx || !g->srcnodeptr || !g->nodenedges || !g->srcnodeidx || !g->dstnodeidx ||
            !g->nbordernodeinneredges || !g->nbordernodeinterfaceedges || !S->a) goto fail;
        for (size_t i = 0; i < nghostp[p]; i++) g->parentnodeidx[no + (acgidx_t) i] = ghosts[p][i].key;
    }
    for (acgidx_t u = 0; u < n; u++) sub[rowparts[u]].graph->parentnodeidx[localof[u]] = u;

    /* count packed entries per local row: an edge (u,v), u<=v in global storage,
     * lands on row local(u) of part(u); a cut edge also on row local(v) of part(v) */
    for (acgidx_t u = 0; u < n; u++) {
        const int pu = rowparts[u];
        for (int64_t k = rp[u]; k < rp[u + 1]; k++) {
            const acgidx_t v = cj[k] - base;
            const int pv = rowparts[v];
            sub[pu].graph->srcnodeptr[localof[u] + 1]++;
            if (pu != pv) sub[pv].graph->srcnodeptr[localof[v] + 1]++;
        }
    }
    for (int p = 0; p < nparts; p++) {
        struct acggraph *g = sub[p].graph;
        for (acgidx_t i = 0; i < g->npnodes; i++) {
            g->nodenedges[i] = g->srcnodeptr[i + 1];
            g->srcnodeptr[i + 1] += g->srcnodeptr[i];
        }
    }
    {
        cur = calloc((size_t) nparts, sizeof(*cur));
        gseg = calloc((size_t) nparts, sizeof(*gseg));   /* start of each owner's ghost segment */
        if (!cur || !gseg) goto fail;
        for (int p = 0; p < nparts; p++) {
            struct acggraph *g = sub[p].graph;
            cur[p] = malloc((size_t) (g->npnodes > 0 ? g->npnodes : 1) * sizeof(int64_t));
            gseg[p] = calloc((size_t) nparts + 1, sizeof(size_t));
            if (!cur[p] || !gseg[p]) goto fail;
            memcpy(cur[p], g->srcnodeptr, (size_t) g->npnodes * sizeof(int64_t));
            for (size_t i = 0; i < nghostp[p]; i++) gseg[p][ghosts[p][i].part + 1]++;
            for (int q = 0; q < nparts; q++) gseg[p][q + 1] += gseg[p][q];
        }
        for (acgidx_t u = 0; u < n; u++) {
            const int pu = rowparts[u];
            for (int64_t k = rp[u]; k < rp[u + 1]; k++) {
                const acgidx_t v = cj[k] - base;
                const int pv = rowparts[v];
                struct acggraph *gu = sub[pu].graph;
                if (pu == pv) {
                    const int64_t l = cur[pu][localof[u]]++;
                    gu->srcnodeidx[l] = localof[u]; gu->dstnodeidx[l] = localof[v];
                    sub[pu].a[l] = A->a[k];
                    gu->ninneredges++;
                    if (isborder[u]) gu->nbordernodeinneredges[localof[u] - gu->bordernodeoffset]++;
                    if (isborder[v] && u != v) gu->nbordernodeinneredges[localof[v] - gu->bordernodeoffset]++;
                } else {
                    struct acggraph *gv = sub[pv].graph;
                    int64_t l = cur[pu][localof[u]]++;
                    gu->srcnodeidx[l] = localof[u];
                    gu->dstnodeidx[l] = gu->ghostnodeoffset + find_pair(ghosts[pu], gseg[pu][pv], gseg[pu][pv + 1], v);
                    sub[pu].a[l] = A->a[k];
                    gu->ninterfaceedges++;
                    gu->nbordernodeinterfaceedges[localof[u] - gu->bordernodeoffset]++;
                    l = cur[pv][localof[v]]++;
                    gv->srcnodeidx[l] = localof[v];
                    gv->dstnodeidx[l] = gv->ghostnodeoffset + find_pair(ghosts[pv], gseg[pv][pu], gseg[pv][pu + 1], u);
                    sub[pv].a[l] = A->a[k];
                    gv->ninterfaceedges++;
                    gv->nbordernodeinterfaceedges[localof[v] - gv->bordernodeoffset]++;
                }
            }
        }
        /* neighbour lists */
        for (int p = 0; p < nparts; p++) {
            struct acggraph *g = sub[p].graph;
            int nn = 0;
            for (int q = 0; q < nparts; q++) if (gseg[p][q + 1] > gseg[p][q]) nn++;
            g->nneighbours = nn;
            g->neighbours = calloc((size_t) (nn > 0 ? nn : 1), sizeof(*g->neighbours));
            if (!g->neighbours) goto fail;
            size_t bpos = 0;
            int i = 0;
            for (int q = 0; q < nparts; q++) {
                if (gseg[p][q + 1] == gseg[p][q]) continue;
                struct acggraphneighbour *nb = &g->neighbours[i++];
                nb->neighbourrank = q; nb->neighbourpart = 0;
                nb->nghostnodes = (acgidx_t) (gseg[p][q + 1] - gseg[p][q]);
                nb->ghostnodes = malloc((size_t) nb->nghostnodes * sizeof(acgidx_t));
                if (!nb->ghostnodes) goto fail;
                for (acgidx_t j = 0; j < nb->nghostnodes; j++) nb->ghostnodes[j] = (acgidx_t) gseg[p][q] + j;
                size_t bend = bpos;
                while (bend < nborderp[p] && borders[p][bend].part == q) bend++;
                nb->nbordernodes = (acgidx_t) (bend - bpos);
                nb->bordernodes = malloc((size_t) (nb->nbordernodes > 0 ? nb->nbordernodes : 1) * sizeof(acgidx_t));
                if (!nb->bordernodes) goto fail;
                for (size_t j = bpos; j < bend; j++)
                    nb->bordernodes[j - bpos] = localof[borders[p][j].key] - g->bordernodeoffset;
                bpos = bend;
            }
            matrix_view_graph(&sub[p]);
            free(cur[p]); free(gseg[p]);
            cur[p] = NULL; gseg[p] = NULL;
        }
    }
    err = ACG_SUCCESS;
fail:
    if (cur) for (int p = 0; p < nparts; p++) free(cur[p]);
    if (gseg) for (int p = 0; p < nparts; p++) free(gseg[p]);
    free(cur); free(gseg);
    if (ghosts) for (int p = 0; p < nparts; p++) free(ghosts[p]);
    if (borders) for (int p = 0; p < nparts; p++) free(borders[p]);
    free(ghosts); free(borders); free(nghostp); free(nborderp);
    free(ncut); free(isborder); free(localof); free(nowned); free(ninner); free(nedgesp);
    if (err) for (int p = 0; p < nparts; p++) if (sub[p].graph) acgsymcsrmatrix_free(&sub[p]);
    return err;
}

/* ------------------------------------------------------------------------ */
/* halo pattern                                                              */
/* ------------------------------------------------------------------------ */

/* acg/graph.c:1898-1981: one send segment and one receive segment per
 * neighbour; send indices address the border range, receive indices the
 * ghost tail of a vector created by acgsymcsrmatrix_vector */
int acgsymcsrmatrix_halo(const struct acgsymcsrmatrix *A, struct acghalo *halo)
{
    const struct acggraph *g = A->graph;
    memset(halo, 0, sizeof(*halo));
    const int nn = g ? g->nneighbours : 0;
    const size_t m = (size_t) (nn > 0 ? nn : 1);
    halo->nrecipients = halo->nsenders = nn;
    halo->recipients = calloc(m, sizeof(int)); halo->sendcounts = calloc(m, sizeof(int)); halo->sdispls = calloc(m, sizeof(int));
    halo->senders = calloc(m, sizeof(int)); halo->recvcounts = calloc(m, sizeof(int)); halo->rdispls = calloc(m, sizeof(int));
    if (!halo->recipients || !halo->sendcounts || !halo->sdispls || !halo->senders || !halo->recvcounts || !halo->rdispls) {
        acghalo_free(halo); return ACG_ERR_ERRNO;
    }
    for (int i = 0; i < nn; i++) {
        const struct acggraphneighbour *nb = &g->neighbours[i];
        halo->recipients[i] = halo->senders[i] = nb->neighbourrank;
        halo->sendcounts[i] = nb->nbordernodes; halo->sdispls[i] = halo->sendsize; halo->sendsize += nb->nbordernodes;
        halo->recvcounts[i] = nb->nghostnodes; halo->rdispls[i] = halo->recvsize; halo->recvsize += nb->nghostnodes;
    }
    halo->sendbufidx = malloc((size_t) (halo->sendsize > 0 ? halo->sendsize : 1) * sizeof(int));
    halo->recvbufidx = malloc((size_t) (halo->recvsize > 0 ? halo->recvsize : 1) * sizeof(int));
    halo->maxexchangestats = 0;
    halo->thaloexchangestats = NULL;
    if (!halo->sendbufidx || !halo->recvbufidx) { acghalo_free(halo); return ACG_ERR_ERRNO; }
    for (int i = 0; i < nn; i++) {
        const struct acggraphneighbour *nb = &g->neighbours[i];
        for (acgidx_t j = 0; j < nb->nbordernodes; j++)
            halo->sendbufidx[halo->sdispls[i] + j] = g->bordernodeoffset + nb->bordernodes[j];
        for (acgidx_t j = 0; j < nb->nghostnodes; j++)
            halo->recvbufidx[halo->rdispls[i] + j] = g->ghostnodeoffset + nb->ghostnodes[j];
    }
    return ACG_SUCCESS;
}

/* ext.h: one row of the communication matrix the reference driver prints with
 * --output-comm-matrix (cuda/acg-cuda.c:1713-1775): entry (p,q) is the number
 * of border values part p sends to part q in one halo exchange.  `row` has
 * nparts entries.  Parts made by this library carry the peer in
 * neighbourrank (the state after the reference's scatter, acg/graph.c:1767-1779). */
int acgb200_comm_matrix_row(const struct acgsymcsrmatrix *A, int nparts, int64_t *row)
{
    const struct acggraph *g = A->graph;
    for (int q = 0; q < nparts; q++) row[q] = 0;
    if (!g) return ACG_SUCCESS;
    for (int i = 0; i < g->nneighbours; i++) {
        const int q = g->neighbours[i].neighbourrank;
        if (q < 0 || q >= nparts) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
        row[q] += g->neighbours[i].nbordernodes;
    }
    return ACG_SUCCESS;
}
