/*
 * stencil.c -- one part of a block-partitioned 7- or 27-point stencil matrix,
 * generated directly (no global matrix, no partitioner).
 *
 * The reference builds distributed problems on rank 0: read or generate the
 * whole matrix, partition it, scatter the parts (cuda/acg-cuda.c:1297-1782).
 * For the synthetic benchmark matrices of BASELINE.json (3-D stencils on an
 * nx*ny*nz box, lexicographic numbering, geometric px*py*pz block partition)
 * every rank can construct its own part analytically in O(local size) time
 * and memory -- which is what makes the 448^3 / 8-GPU configuration
 * (2.4e9 nonzeros) fit without any process ever holding the global matrix.
 *
 * The result is array-for-array what
 *     acgsymcsrmatrix_partition(global stencil matrix, block row->part map)[part]
 * produces (tests/test_host_structs.py::test_stencil_part_matches_partition):
 * local order [interior | border | ghost], ghosts grouped by owner and
 * ascending in global index, packed rows holding first the cut edges towards
 * lower-numbered foreign rows, then the row's own upper-triangle entries.
 */
#include "acgb200/error.h"
#include "acgb200/ext.h"
#include "acgb200/symcsrmatrix.h"
#include "hostmem.h"

#include <stdlib.h>
#include <string.h>

static inline int blk_lo(int i, int n, int p) { return (int) (((int64_t) i * n + p - 1) / p); }   /* first x with x*p/n == i */
static inline int blk_of(int x, int n, int p) { return (int) ((int64_t) x * p / n); }

struct geom {
    int kind, nx, ny, nz, px, py, pz;
    int x0, x1, y0, y1, z0, z1;     /* this part's box */
    int part;
};

static inline int owner_of(const struct geom *g, int x, int y, int z)
{
    return blk_of(x, g->nx, g->px) + g->px * (blk_of(y, g->ny, g->py) + g->py * blk_of(z, g->nz, g->pz));
}
static inline int in_grid(const struct geom *g, int x, int y, int z)
{
    return x >= 0 && x < g->nx && y >= 0 && y < g->ny && z >= 0 && z < g->nz;
}
static inline int in_box(const struct geom *g, int x, int y, int z)
{
    return x >= g->x0 && x < g->x1 && y >= g->y0 && y < g->y1 && z >= g->z0 && z < g->z1;
}
static inline int is_stencil(const struct geom *g, int dx, int dy, int dz)
{
    return g->kind == 27 ? 1 : (abs(dx) + abs(dy) + abs(dz) <= 1);
}
static inline int64_t gidx(const struct geom *g, int x, int y, int z)
{
    return x + (int64_t) g->nx * (y + (int64_t) g->ny * z);
}

struct gpair { int64_t idx; int owner; };
static int cmp_gpair(const void *a, const void *b)
{
    const struct gpair *x = a, *y = b;
    if (x->owner != y->owner) return x->owner < y->owner ? -1 : 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

int acgb200_stencil_part(int kind, int nx, int ny, int nz, int px, int py, int pz, int part,
                         struct acgsymcsrmatrix *A)
{
    if ((kind != 7 && kind != 27) || nx < 3 || ny < 3 || nz < 3 || px < 1 || py < 1 || pz < 1 ||
        px > nx || py > ny || pz > nz || part < 0 || part >= px * py * pz) return ACG_ERR_INVALID_VALUE;
    if ((int64_t) nx * ny * nz > ACGIDX_T_MAX) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    memset(A, 0, sizeof(*A));
    struct geom g = { kind, nx, ny, nz, px, py, pz, 0, 0, 0, 0, 0, 0, part };
    const int bi = part % px, bj = (part / px) % py, bk = part / (px * py);
    g.x0 = blk_lo(bi, nx, px); g.x1 = blk_lo(bi + 1, nx, px);
    g.y0 = blk_lo(bj, ny, py); g.y1 = blk_lo(bj + 1, ny, py);
    g.z0 = blk_lo(bk, nz, pz); g.z1 = blk_lo(bk + 1, nz, pz);
    const int bx = g.x1 - g.x0, by = g.y1 - g.y0, bz = g.z1 - g.z0;
    const int64_t nown = (int64_t) bx * by * bz;
    const double diag = kind == 27 ? 26.0 : 6.0;
    int err = ACG_ERR_ERRNO;

    /* local numbers of the box (owned) and of its one-cell shell (ghost candidates) */
    const int sx = bx + 2, sy = by + 2, sz = bz + 2;
    int *loc = acgb200_bigalloc((size_t) (nown > 0 ? nown : 1) * sizeof(int));
    int *shell = acgb200_bigalloc((size_t) sx * sy * sz * sizeof(int));
    unsigned char *isb = calloc((size_t) (nown > 0 ? nown : 1), 1);
    struct gpair *gh = NULL;
    struct acggraph *gr = calloc(1, sizeof(*gr));
    if (!loc || !shell || !isb || !gr) goto fail;
#define BOX(x, y, z) ((size_t) ((z) - g.z0) * by * bx + (size_t) ((y) - g.y0) * bx + (size_t) ((x) - g.x0))
#define SHELL(x, y, z) ((size_t) ((z) - g.z0 + 1) * sy * sx + (size_t) ((y) - g.y0 + 1) * sx + (size_t) ((x) - g.x0 + 1))

    /* classify owned nodes; collect ghosts */
    int64_t ninner = 0, nghost_cap = (int64_t) sx * sy * sz - nown, ngh = 0;
    gh = malloc((size_t) (nghost_cap > 0 ? nghost_cap : 1) * sizeof(*gh));
    if (!gh) goto fail;
    for (size_t i = 0; i < (size_t) sx * sy * sz; i++) shell[i] = -1;
    for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++) {
        int border = 0;
        if (x > g.x0 && x < g.x1 - 1 && y > g.y0 && y < g.y1 - 1 && z > g.z0 && z < g.z1 - 1) {
            ninner++;                            /* off the faces of the box: interior */
            continue;
        }
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            if (!is_stencil(&g, dx, dy, dz)) continue;
            const int X = x + dx, Y = y + dy, Z = z + dz;
            if (!in_grid(&g, X, Y, Z) || in_box(&g, X, Y, Z)) continue;
            border = 1;
            if (shell[SHELL(X, Y, Z)] == -1) {
                shell[SHELL(X, Y, Z)] = -2;      /* seen */
                gh[ngh].idx = gidx(&g, X, Y, Z); gh[ngh].owner = owner_of(&g, X, Y, Z); ngh++;
            }
        }
        isb[BOX(x, y, z)] = (unsigned char) border;
        ninner += !border;
    }
    const int64_t nborder = nown - ninner;
    {   /* local order: interior ascending, then border ascending */
        int ni = 0, nb = 0;
        for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++)
            loc[BOX(x, y, z)] = isb[BOX(x, y, z)] ? (int) ninner + nb++ : ni++;
    }
    qsort(gh, (size_t) ngh, sizeof(*gh), cmp_gpair);
    for (int64_t i = 0; i < ngh; i++) {
        const int64_t id = gh[i].idx;
        const int X = (int) (id % nx), Y = (int) ((id / nx) % ny), Z = (int) (id / ((int64_t) nx * ny));
        shell[SHELL(X, Y, Z)] = (int) (nown + i);
    }
    const int64_t npn = nown + ngh;

    /* graph skeleton */
    gr->nparts = px * py * pz; gr->parttag = part + 1; gr->nprocs = gr->nparts; gr->npparts = 1;
    gr->ownerrank = part; gr->ownerpart = 0;
    gr->nnodes = (acgidx_t) ((int64_t) nx * ny * nz); gr->npnodes = (acgidx_t) npn;
    {
        const int64_t n = (int64_t) nx * ny * nz;
        const int64_t full = kind == 27 ? (int64_t) (3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2)
            : n + 2 * ((int64_t) (nx - 1) * ny * nz + (int64_t) nx * (ny - 1) * nz + (int64_t) nx * ny * (nz - 1));
        gr->nedges = (full + n) / 2;
    }
    gr->nodeidxbase = 0;
    gr->nownednodes = (acgidx_t) nown; gr->ninnernodes = (acgidx_t) ninner; gr->nbordernodes = (acgidx_t) nborder;
    gr->bordernodeoffset = (acgidx_t) ninner; gr->nghostnodes = (acgidx_t) ngh; gr->ghostnodeoffset = (acgidx_t) nown;
    gr->parentnodeidx = malloc((size_t) (npn > 0 ? npn : 1) * sizeof(acgidx_t));
    gr->srcnodeptr = calloc((size_t) npn + 1, sizeof(int64_t));
    gr->nodenedges = calloc((size_t) (npn > 0 ? npn : 1), sizeof(int64_t));
    gr->nbordernodeinneredges = calloc((size_t) (nborder > 0 ? nborder : 1), sizeof(int64_t));
    gr->nbordernodeinterfaceedges = calloc((size_t) (nborder > 0 ? nborder : 1), sizeof(int64_t));
    if (!gr->parentnodeidx || !gr->srcnodeptr || !gr->nodenedges || !gr->nbordernodeinneredges ||
        !gr->nbordernodeinterfaceedges) goto fail;
    for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++)
        gr->parentnodeidx[loc[BOX(x, y, z)]] = (acgidx_t) gidx(&g, x, y, z);
    for (int64_t i = 0; i < ngh; i++) gr->parentnodeidx[nown + i] = (acgidx_t) gh[i].idx;

    /* packed rows: count, then fill.  Row of owned node u holds
     *   (a) cut edges to foreign neighbours with a smaller global index, ascending,
     *   (b) its own upper-triangle entries (global index >= u), ascending,
     * which is the order acgsymcsrmatrix_partition produces. */
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (int64_t i = 0; i < npn; i++) { gr->nodenedges[i] = gr->srcnodeptr[i + 1]; gr->srcnodeptr[i + 1] += gr->srcnodeptr[i]; }
            gr->npedges = gr->srcnodeptr[npn];
            const size_t ne = (size_t) (gr->npedges > 0 ? gr->npedges : 1);
            gr->srcnodeidx = acgb200_bigalloc(ne * sizeof(acgidx_t));
            gr->dstnodeidx = acgb200_bigalloc(ne * sizeof(acgidx_t));
            A->a = acgb200_bigalloc(ne * sizeof(double));
            if (!gr->srcnodeidx || !gr->dstnodeidx || !A->a) goto fail;
        }
        #pragma omp parallel for collapse(2)
        for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++) {
            const int lu = loc[BOX(x, y, z)];
            int64_t pos = pass == 1 ? gr->srcnodeptr[lu] : 0;
            for (int half = 0; half < 2; half++) {
                for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
                    if (!is_stencil(&g, dx, dy, dz)) continue;
                    const int lower = dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx < 0)));
                    if (lower != (half == 0)) continue;
                    const int X = x + dx, Y = y + dy, Z = z + dz;
                    if (!in_grid(&g, X, Y, Z)) continue;
                    const int inside = in_box(&g, X, Y, Z);
                    if (lower && inside) continue;          /* stored at the lower row */
                    if (pass == 0) { gr->srcnodeptr[lu + 1]++; continue; }
                    gr->srcnodeidx[pos] = lu;
                    gr->dstnodeidx[pos] = inside ? loc[BOX(X, Y, Z)] : shell[SHELL(X, Y, Z)];
                    A->a[pos] = (dx == 0 && dy == 0 && dz == 0) ? diag : -1.0;
                    pos++;
                }
            }
        }
    }
    /* edge statistics */
    int64_t ninneredges = 0, ninterfaceedges = 0;
    #pragma omp parallel for collapse(2) reduction(+:ninneredges, ninterfaceedges)
    for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++) {
        const int lu = loc[BOX(x, y, z)];
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            if (!is_stencil(&g, dx, dy, dz)) continue;
            const int X = x + dx, Y = y + dy, Z = z + dz;
            if (!in_grid(&g, X, Y, Z)) continue;
            const int lower = dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx < 0)));
            if (in_box(&g, X, Y, Z)) {
                if (lower) continue;
                ninneredges++;
                if (isb[BOX(x, y, z)])
                    __atomic_fetch_add(&gr->nbordernodeinneredges[lu - ninner], 1, __ATOMIC_RELAXED);
                if (isb[BOX(X, Y, Z)] && !(dx == 0 && dy == 0 && dz == 0))
                    __atomic_fetch_add(&gr->nbordernodeinneredges[loc[BOX(X, Y, Z)] - ninner], 1, __ATOMIC_RELAXED);
            } else {
                ninterfaceedges++;
                __atomic_fetch_add(&gr->nbordernodeinterfaceedges[lu - ninner], 1, __ATOMIC_RELAXED);
            }
        }
    }
    gr->ninneredges = ninneredges; gr->ninterfaceedges = ninterfaceedges;
    /* neighbours: one per foreign owner among the ghosts, ascending */
    {
        int nn = 0;
        for (int64_t i = 0; i < ngh; i++) if (i == 0 || gh[i].owner != gh[i - 1].owner) nn++;
        gr->nneighbours = nn;
        gr->neighbours = calloc((size_t) (nn > 0 ? nn : 1), sizeof(*gr->neighbours));
        if (!gr->neighbours) goto fail;
        int64_t i = 0;
        for (int q = 0; q < nn; q++) {
            struct acggraphneighbour *nb = &gr->neighbours[q];
            const int owner = gh[i].owner;
            const int64_t first = i;
            while (i < ngh && gh[i].owner == owner) i++;
            nb->neighbourrank = owner; nb->neighbourpart = 0;
            nb->nghostnodes = (acgidx_t) (i - first);
            nb->ghostnodes = malloc((size_t) nb->nghostnodes * sizeof(acgidx_t));
            if (!nb->ghostnodes) goto fail;
            for (int64_t j = first; j < i; j++) nb->ghostnodes[j - first] = (acgidx_t) j;
            /* border nodes adjacent to that owner, ascending (= ascending local index) */
            acgidx_t cnt = 0;
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) {
                    nb->nbordernodes = cnt;
                    nb->bordernodes = malloc((size_t) (cnt > 0 ? cnt : 1) * sizeof(acgidx_t));
                    if (!nb->bordernodes) goto fail;
                    cnt = 0;
                }
                for (int z = g.z0; z < g.z1; z++) for (int y = g.y0; y < g.y1; y++) for (int x = g.x0; x < g.x1; x++) {
                    if (!isb[BOX(x, y, z)]) continue;
                    int adj = 0;
                    for (int dz = -1; dz <= 1 && !adj; dz++) for (int dy = -1; dy <= 1 && !adj; dy++) for (int dx = -1; dx <= 1 && !adj; dx++) {
                        if (!is_stencil(&g, dx, dy, dz)) continue;
                        const int X = x + dx, Y = y + dy, Z = z + dz;
                        if (in_grid(&g, X, Y, Z) && !in_box(&g, X, Y, Z) && owner_of(&g, X, Y, Z) == owner) adj = 1;
                    }
                    if (!adj) continue;
                    if (pass == 1) nb->bordernodes[cnt] = (acgidx_t) (loc[BOX(x, y, z)] - ninner);
                    cnt++;
                }
            }
        }
    }
#undef BOX
#undef SHELL
    /* matrix views of the graph (as symcsrmatrix.c does for partitioned parts) */
    {
        double *vals = A->a;
        memset(A, 0, sizeof(*A));
        A->a = vals;
        A->graph = gr;
        A->nrows = gr->nnodes; A->nprows = gr->npnodes; A->nzrows = gr->parentnodeidx;
        A->nnzs = gr->nedges; A->npnzs = gr->npedges; A->rowidxbase = 0;
        A->rownnzs = gr->nodenedges; A->rowptr = gr->srcnodeptr; A->rowidx = gr->srcnodeidx; A->colidx = gr->dstnodeidx;
        A->nownedrows = gr->nownednodes; A->ninnerrows = gr->ninnernodes; A->nborderrows = gr->nbordernodes;
        A->borderrowoffset = gr->bordernodeoffset; A->nghostrows = gr->nghostnodes; A->ghostrowoffset = gr->ghostnodeoffset;
        A->ninnernzs = gr->ninneredges; A->ninterfacenzs = gr->ninterfaceedges;
        A->nborderrowinnernzs = gr->nbordernodeinneredges; A->nborderrowinterfacenzs = gr->nbordernodeinterfaceedges;
    }
    free(loc); free(shell); free(isb); free(gh);
    return ACG_SUCCESS;
fail:
    free(loc); free(shell); free(isb); free(gh);
    if (gr) {
        struct acgsymcsrmatrix tmp;
        memset(&tmp, 0, sizeof(tmp));
        tmp.graph = gr; tmp.a = A->a;
        acgsymcsrmatrix_free(&tmp);
    }
    return err;
}

/* ext.h: geometric row -> part map of an nx*ny*nz lexicographic grid cut into
 * px*py*pz blocks (part = bi + px*(bj + py*bk)) -- the partition
 * acgb200_stencil_part assumes, usable with acgsymcsrmatrix_partition or
 * acgb200_mtx_read_part for any matrix on such a grid; the alternative to the
 * METIS call in acg/graph.c:510 when the geometry is known. */
int acgb200_partition_rows_grid(int nx, int ny, int nz, int px, int py, int pz, int *rowparts)
{
    if (nx < 1 || ny < 1 || nz < 1 || px < 1 || py < 1 || pz < 1 || px > nx || py > ny || pz > nz)
        return ACG_ERR_INVALID_VALUE;
    #pragma omp parallel for collapse(2)
    for (int z = 0; z < nz; z++) for (int y = 0; y < ny; y++) {
        const int base = px * (blk_of(y, ny, py) + py * blk_of(z, nz, pz));
        int *row = rowparts + ((size_t) z * ny + y) * nx;
        for (int x = 0; x < nx; x++) row[x] = blk_of(x, nx, px) + base;
    }
    return ACG_SUCCESS;
}

/* ext.h: px*py*pz = nparts, as cubic as possible, px <= py <= pz */
void acgb200_grid_factors(int nparts, int *px, int *py, int *pz)
{
    int best[3] = { 1, 1, nparts };
    for (int a = 1; a <= nparts; a++) {
        if (nparts % a) continue;
        for (int b = a; b <= nparts / a; b++) {
            if ((nparts / a) % b) continue;
            const int c = nparts / a / b;
            if (c >= b && (c - a) < (best[2] - best[0])) { best[0] = a; best[1] = b; best[2] = c; }
        }
    }
    *px = best[0]; *py = best[1]; *pz = best[2];
}
