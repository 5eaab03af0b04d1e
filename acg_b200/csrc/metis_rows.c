/*
 * metis_rows.c -- METIS row partitioning of a symmetric CSR matrix.
 *
 * Own implementation of acgsymcsrmatrix_partition_rows (reference:
 * acg/symcsrmatrix.c:656-683 -> acg/graph.c:510 -> metis_partgraphsym,
 * acg/metis.c:80-382): the sparsity pattern of the packed upper triangle is
 * turned into an unweighted, undirected adjacency structure without self
 * loops and handed to METIS_PartGraphRecursive / METIS_PartGraphKway.
 *
 * METIS itself is the static library that ships inside the CUDA toolkit
 * (targets/x86_64-linux/lib/libmetis_static.a, METIS 5.x built with 64-bit
 * idx_t and 32-bit real_t); it has no header there, hence the prototypes
 * below.  Without that archive the function returns
 * ACG_ERR_METIS_NOT_SUPPORTED, like the reference built without METIS.
 */
#include "acgb200/error.h"
#include "acgb200/symcsrmatrix.h"

#include <stdint.h>
#include <stdlib.h>

#ifdef ACG_HAVE_METIS
typedef int64_t idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
#define METIS_OPTION_SEED 8
#define METIS_OK 1
#define METIS_ERROR_INPUT (-2)
#define METIS_ERROR_MEMORY (-3)
int METIS_SetDefaultOptions(idx_t *options);
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
                             idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
                             idx_t *edgecut, idx_t *part);
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
                        idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
                        idx_t *edgecut, idx_t *part);
#endif

int acgsymcsrmatrix_partition_rows(
    struct acgsymcsrmatrix *A, int nparts, enum metis_partitioner partitioner,
    int *rowparts, acgidx_t *objval, acgidx_t seed, int verbose)
{
    (void) verbose;
    const acgidx_t n = A->nprows;
    if (nparts < 1) return ACG_ERR_INVALID_VALUE;
    if (nparts == 1 || n == 0) {
        for (acgidx_t i = 0; i < n; i++) rowparts[i] = 0;
        if (objval) *objval = 0;
        return ACG_SUCCESS;
    }
#ifndef ACG_HAVE_METIS
    (void) partitioner; (void) seed;
    return ACG_ERR_METIS_NOT_SUPPORTED;
#else
    const int base = A->rowidxbase;
    const int64_t *rp = A->rowptr;
    const acgidx_t *cj = A->colidx;
    idx_t *xadj = calloc((size_t) n + 2, sizeof(*xadj));
    if (!xadj) return ACG_ERR_ERRNO;
    for (acgidx_t i = 0; i < n; i++) {
        for (int64_t k = rp[i]; k < rp[i + 1]; k++) {
            const acgidx_t j = cj[k] - base;
            if (j != i) { xadj[i + 2]++; xadj[j + 2]++; }
        }
    }
    for (acgidx_t i = 0; i < n; i++) xadj[i + 2] += xadj[i + 1];
    const idx_t nadj = xadj[n + 1];
    idx_t *adj = malloc((size_t) (nadj > 0 ? nadj : 1) * sizeof(*adj));
    idx_t *part = malloc((size_t) n * sizeof(*part));
    if (!adj || !part) { free(xadj); free(adj); free(part); return ACG_ERR_ERRNO; }
    for (acgidx_t i = 0; i < n; i++) {
        for (int64_t k = rp[i]; k < rp[i + 1]; k++) {
            const acgidx_t j = cj[k] - base;
            if (j != i) { adj[xadj[i + 1]++] = j; adj[xadj[j + 1]++] = i; }
        }
    }
    /* xadj[i+1] now holds the end of row i, i.e. xadj[0..n] is the usual pointer array */
    idx_t nv = n, ncon = 1, np = nparts, cut = 0, options[METIS_NOPTIONS];
    METIS_SetDefaultOptions(options);
    options[METIS_OPTION_SEED] = seed;
    int merr = partitioner == metis_partgraphkway
        ? METIS_PartGraphKway(&nv, &ncon, xadj, adj, NULL, NULL, NULL, &np, NULL, NULL, options, &cut, part)
        : METIS_PartGraphRecursive(&nv, &ncon, xadj, adj, NULL, NULL, NULL, &np, NULL, NULL, options, &cut, part);
    int err = ACG_SUCCESS;
    if (merr == METIS_ERROR_INPUT) err = ACG_ERR_METIS_INPUT;
    else if (merr == METIS_ERROR_MEMORY) err = ACG_ERR_METIS_MEMORY;
    else if (merr != METIS_OK) err = ACG_ERR_METIS;
    if (!err) {
        for (acgidx_t i = 0; i < n; i++) rowparts[i] = (int) part[i];
        if (objval) *objval = (acgidx_t) cut;
    }
    free(xadj); free(adj); free(part);
    return err;
#endif
}
