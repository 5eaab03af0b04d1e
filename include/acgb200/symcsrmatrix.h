/*
 * acgb200/symcsrmatrix.h -- symmetric CSR matrices, row-partitioned.
 *
 * ABI counterpart of acg/symcsrmatrix.h:62-292 (struct acgsymcsrmatrix) and
 * acg/graph.h:54-329 (struct acggraph, which the matrix points to and whose
 * neighbour lists define the halo pattern).  The device solver consumes only
 * the "full storage" arrays
 *
 *     frowptr/fcolidx/fa    local block      rows [0,nownedrows)
 *     orowptr/ocolidx/oa    border x ghost   rows [borderrowoffset, nprows),
 *                           column indices rebased by -borderrowoffset
 *
 * built by acgsymcsrmatrix_dsymv_init (acg/symcsrmatrix.c:760-851), plus the
 * row-class counters and graph->neighbours.  Local row order is
 * [interior | border | ghost] (acg/graph.c:813).
 */
#ifndef ACGB200_SYMCSRMATRIX_H
#define ACGB200_SYMCSRMATRIX_H

#include "acgb200/config.h"
#include "acgb200/vector.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct acghalo;

/* acg/graph.h:276-321 */
struct acggraphneighbour {
    int neighbourrank;
    int neighbourpart;
    acgidx_t nbordernodes;
    acgidx_t *bordernodes;   /* relative to bordernodeoffset */
    acgidx_t nghostnodes;
    acgidx_t *ghostnodes;    /* relative to ghostnodeoffset */
};

/* acg/graph.h:54-329 */
struct acggraph {
    int nparts, parttag, nprocs, npparts, ownerrank, ownerpart;
    acgidx_t nnodes, npnodes;
    acgidx_t *nodetags;
    acgidx_t *parentnodeidx;
    int64_t nedges, npedges;
    int64_t *edgetags;
    int64_t *parentedgeidx;
    int nodeidxbase;
    int64_t *nodenedges;
    int64_t *srcnodeptr;
    acgidx_t *srcnodeidx;
    acgidx_t *dstnodeidx;
    acgidx_t nownednodes, ninnernodes, nbordernodes, bordernodeoffset;
    acgidx_t nghostnodes, ghostnodeoffset;
    int64_t ninneredges, ninterfaceedges;
    int64_t *nbordernodeinneredges;
    int64_t *nbordernodeinterfaceedges;
    int nneighbours;
    struct acggraphneighbour *neighbours;
};

/* acg/symcsrmatrix.h:62-292; the const members alias arrays owned by *graph */
struct acgsymcsrmatrix {
    struct acggraph *graph;
    acgidx_t nrows, nprows;
    const acgidx_t *nzrows;          /* global row number of each local row (NULL if unpartitioned) */
    int64_t nnzs, npnzs;
    int rowidxbase;
    const int64_t *rownnzs;
    const int64_t *rowptr;           /* packed upper triangle */
    const acgidx_t *rowidx;
    const acgidx_t *colidx;
    acgidx_t nownedrows, ninnerrows, nborderrows, borderrowoffset;
    acgidx_t nghostrows, ghostrowoffset;
    int64_t ninnernzs, ninterfacenzs;
    const int64_t *nborderrowinnernzs;
    const int64_t *nborderrowinterfacenzs;
    double *a;
    int64_t fnpnzs, onpnzs;          /* full storage, see file comment */
    int64_t *frowptr, *orowptr;
    acgidx_t *fcolidx, *ocolidx;
    double *fa, *oa;
};

/* acg/symcsrmatrix.h:301 -- COO upper triangle -> packed CSR */
ACG_API int acgsymcsrmatrix_init_real_double(
    struct acgsymcsrmatrix *A, acgidx_t N, int64_t nnzs, int idxbase,
    const acgidx_t *rowidx, const acgidx_t *colidx, const double *a);
/* acg/symcsrmatrix.h:317 */
ACG_API int acgsymcsrmatrix_init_rowwise_real_double(
    struct acgsymcsrmatrix *A, acgidx_t N, int idxbase,
    const int64_t *rowptr, const acgidx_t *colidx, const double *a);
/* acg/symcsrmatrix.h:328 */
ACG_API void acgsymcsrmatrix_free(struct acgsymcsrmatrix *A);
/* acg/symcsrmatrix.h:405 -- vector compatible with A (owned + ghost entries) */
ACG_API int acgsymcsrmatrix_vector(const struct acgsymcsrmatrix *A, struct acgvector *x);
/* acg/metis.h:39-43 */
enum metis_partitioner { metis_partgraphrecursive, metis_partgraphkway };
/* acg/symcsrmatrix.h:419 -- row->part map (0-based part numbers) from METIS on
 * the matrix graph; objval receives the edge cut */
ACG_API int acgsymcsrmatrix_partition_rows(
    struct acgsymcsrmatrix *A, int nparts, enum metis_partitioner partitioner,
    int *rowparts, acgidx_t *objval, acgidx_t seed, int verbose);
/* acg/symcsrmatrix.h:435 -- split by a row->part map into nparts submatrices
 * with [interior|border|ghost] local order and neighbour lists */
ACG_API int acgsymcsrmatrix_partition(
    const struct acgsymcsrmatrix *A, int nparts, const int *rowparts,
    struct acgsymcsrmatrix *submatrices, int verbose);
/* acg/symcsrmatrix.h:495 */
ACG_API int acgsymcsrmatrix_halo(const struct acgsymcsrmatrix *A, struct acghalo *halo);
/* acg/symcsrmatrix.h:503 -- packed -> full local CSR + border x ghost CSR */
ACG_API int acgsymcsrmatrix_dsymv_init(struct acgsymcsrmatrix *A, double eps);

#ifdef __cplusplus
}
#endif
#endif
