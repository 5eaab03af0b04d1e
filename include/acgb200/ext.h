/*
 * acgb200/ext.h -- entry points that have no counterpart in the reference's
 * headers.  They exist for tests, bench.py and bindings (ctypes cannot size
 * the ABI structs by itself); nothing in the reference driver needs them.
 */
#ifndef ACGB200_EXT_H
#define ACGB200_EXT_H

#include "acgb200/cgcuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* tunables: "profile" (0/1: CUDA-event timing of each kernel class, fills
 * tgemv/taxpy like the reference's -DACG_ENABLE_PROFILING, acg/cgcuda.c:69-73),
 * "check_every" (iterations between convergence polls), "spmv_lanes",
 * "spmv_nnz_cap", "spmv_rows_cap", "spmv_stages", "spmv_threads", "spmv_unroll",
 * "spmv_max_ctas" (SpMV tile plan overrides, read at acgsolvercuda_init), "spmv_slices"
 * (0/1, default 1: index-free slice-major storage of the rows that repeat a pattern, slices.c; shape
 * overrides "slice_ub", "slice_threads", "slice_max_ctas"), "spmv_merge" (-1/0/1,
 * default -1: merge-path tiles when the row lengths are irregular, mergeplan.c; "merge_items", "merge_threads",
 * "merge_max_ctas"), "spmv_medium", "graph"
 * (0/1: replay iteration pairs as CUDA graphs), "redstream" (0/1: pipelined
 * allreduce on its own stream and communicator; read at acgsolvercuda_init).  Environment variables ACGB200_<KEY> set the
 * same values at first use. */
ACG_API int acgb200_set_option(const char *key, int value);

/* y = A*x through the solver's device matrix: the standalone form of the
 * kernel that replaces cusparseSpMV (acg/cgcuda.c:858).  x, y are HOST arrays
 * of nownedrows doubles (unpartitioned matrices only).  If nrep > 0 the kernel
 * is additionally launched nrep times between two CUDA events on its own
 * stream and the mean duration is returned in milliseconds. */
ACG_API int acgsolvercuda_spmv(struct acgsolvercuda *cg, const double *x, double *y, int nrep, double *ms_per_spmv);

/* One part's share of y = A x on one device with the ghost entries of x supplied by the caller:
 * x has num_nonzeros = owned + ghost entries (host), y the owned rows (host).  path 0: local block +
 * offdiag_kernel on the ghost tail of x (set-up products / NCCL loop, acg/cgcuda.c:858,:878); path 1:
 * the fused tile kernel of the peer-memory loop reading ghosts from a (loop-back) window after the
 * flag wait.  *dot (may be NULL) = sum over owned rows of x_i*y_i from the fused epilogue.  For
 * single-GPU parity tests of the border x ghost path of partitioned matrices. */
ACG_API int acgsolvercuda_spmv_ghost(struct acgsolvercuda *cg, const double *x, double *y, int path, double *dot);

struct acgb200_info {
    int spmv_lanes_per_row, spmv_rows_cap, spmv_nnz_cap, spmv_stages;
    int spmv_ntiles, spmv_nlong, spmv_grid, spmv_smem_bytes;
    int num_sms;
    int last_launches;          /* kernels + NCCL launches enqueued by the last solve's timed region */
    int last_spmv_count;        /* SpMV applications timed in the last solve (profile=1) */
    double last_spmv_ms;        /* their total device time, CUDA events on the launching stream */
    double last_solve_ms;       /* device time of the last solve window (after the H2D copies and warm-up, to the
                                   end of the loop: the window of tsolve, acg/cgcuda.c:719-722,:1021), CUDA events */
    double last_h2d_ms;         /* host time of the b, x0 upload of the last solve */
    double last_d2h_ms;         /* host time of the x download of the last solve */
    double last_blas_ms;        /* device time of the fused vector-update kernels of the last solve (profile=1) */
    int reserved0;              /* (was: tiles without column indices, a round-1 variant removed after measurement) */
    int64_t spmv_min_bytes;     /* bytes one SpMV launch must move at least, given the plan */
    int spmv_nmedium;           /* rows handled one warp each (option "spmv_medium") */
    int reserved1;              /* (round 1: loop layout, the variants were removed) */
    int spmv_slices;            /* 32-row pattern slices multiplied by spmv_slices_kernel (option "spmv_slices", slices.c) */
    int spmv_slice_rows;        /* rows they cover; the other rows are in tiles */
    int spmv_slice_ub, spmv_slice_grid;
    int spmv_merge_tiles;       /* merge-path tiles (option "spmv_merge", mergeplan.c) ... */
    int spmv_merge_rows;        /* ... the rows [0, spmv_merge_rows) they cover ... */
    int spmv_merge_split;       /* ... and the rows cut by tile boundaries (finished by spmv_merge_fix_kernel) */
    int spmv_slice_exc;         /* rows inside slices that are not in the pattern dictionary (columns from the index array) */
};
ACG_API int acgsolvercuda_info(const struct acgsolvercuda *cg, struct acgb200_info *info);

/* The SpMV tile plan acgsolvercuda_init would build for a CSR row-pointer array
 * (int64, nrows+1 entries), computed on the host without a device: tiles4 gets
 * {row_begin, nrows, k_al, nnz_al} per tile, longrows the rows that exceed a
 * tile; info->spmv_* describe the plan.  For tests of the host logic. */
ACG_API int acgb200_spmv_plan_host(int nrows, const int64_t *rowptr, struct acgb200_info *info,
                                   int *tiles4, int maxtiles, int *longrows, int maxlong);
/* same, and with colidx != NULL (0-based) also the pattern slices: the tiles then skip the rows of
 * covered slices; info->spmv_slices / spmv_slice_rows count them (acgb200_slices_host lists them) */
ACG_API int acgb200_spmv_plan_host2(int nrows, const int64_t *rowptr, const int *colidx, struct acgb200_info *info,
                                    int *tiles4, int maxtiles, int *longrows, int maxlong);

/* The pattern-slice plan of a 0-based CSR matrix (slices.c), host only: slices4 gets {row0, nrows,
 * len, vblk} per covered slice (at most maxslices), covered one byte per 32-row slice, totals
 * {nslices, value blocks of 32 doubles, nonzeros covered, rows covered, row stride of the padded
 * offset table, patterns in the table}, spatoff (8192 ints) the zero-padded offset table, patid (nrows, may be
 * NULL) the pattern id per row as the slice kernel sees it (0xFFFF: exception row, columns from the index array),
 * exc2 (may be NULL) {exception rows inside slices, their nonzeros}.  Rows [0,cover_hi) are eligible. */
ACG_API int acgb200_slices_host(int nrows, int cover_hi, const int64_t *rowptr, const int *colidx,
                                int *slices4, int maxslices, unsigned char *covered, int64_t *totals6, int *spatoff,
                                unsigned short *patid, int64_t *exc2);

/* acgsymcsrmatrix_dsymv_init (acg/symcsrmatrix.c:760-851) computed on the current CUDA device
 * (expand.cu): the packed triangle is uploaded, mirrored there into the full local block and the
 * border x ghost block, and copied back into A -- arrays byte-identical to the host routine's.
 * ACG_ERR_CUDA without a device (the host routine is the one to call then).  Programs that only
 * want to solve need not call either: acgsolvercuda_init expands on the device by itself when the
 * matrix has no full storage (diagonal shift 0), which also halves the matrix upload. */
ACG_API int acgsymcsrmatrix_dsymv_init_cuda(struct acgsymcsrmatrix *A, double eps, int *cudaerrcode);

/* The merge-path tile plan of rows [0,hi) of a CSR row-pointer array (mergeplan.c), host only: tiles4 gets
 * {r0, nre, k0, nnz} per tile (at most maxtiles), split3 {row, ta, tb} per row cut by tile boundaries (at
 * most maxtiles); *ntiles, *nsplit the counts. */
ACG_API int acgb200_merge_plan_host(int hi, const int64_t *rowptr, int items, int *tiles4, int maxtiles,
                                    int *split3, int *ntiles, int *nsplit);

/* One part of the block-partitioned 7- or 27-point stencil matrix on an
 * nx*ny*nz box (diag 6 / 26, neighbours -1, lexicographic numbering, px*py*pz
 * geometric blocks, part = bi + px*(bj + py*bk)), built directly without a
 * global matrix: array-for-array what acgsymcsrmatrix_partition yields for that
 * part (stencil.c).  Follow with acgsymcsrmatrix_dsymv_init. */
ACG_API int acgb200_stencil_part(int kind, int nx, int ny, int nz, int px, int py, int pz, int part,
                                 struct acgsymcsrmatrix *A);

/* R-MAT style power-law SPD matrix (BASELINE config 5): n vertices, `nedges`
 * draws with quadrant probabilities abcd (NULL: 0.57, 0.19, 0.19, 0.05),
 * symmetrised, deduplicated, off-diagonal -1, diagonal degree+1.  Threaded; the
 * result does not depend on the thread count (rmat.c). */
ACG_API int acgb200_rmat_spd(int64_t n, int64_t nedges, uint64_t seed, const double *abcd,
                             struct acgsymcsrmatrix *A);

/* Geometric row -> part map of an nx*ny*nz lexicographic grid in px*py*pz blocks
 * (the partition acgb200_stencil_part assumes; the alternative to the METIS call
 * of acg/graph.c:510 when the geometry is known), and the most cubic
 * factorisation px <= py <= pz of a part count. */
ACG_API int acgb200_partition_rows_grid(int nx, int ny, int nz, int px, int py, int pz, int *rowparts);
ACG_API void acgb200_grid_factors(int nparts, int *px, int *py, int *pz);

/* Matrix Market ingest (mtxfile.c): text, and aCG's binary encoding -- header and
 * size line as text, then rowidx[nnz], colidx[nnz] (1-based acgidx_t), a[nnz]
 * (double); acg/mtxfile.c:1107-1127, written by mtx2bin (mtx2bin/mtx2bin.c:538-549),
 * read by `acg-cuda --binary` (cuda/acg-cuda.c:1297-1304).  Only "matrix
 * coordinate real symmetric" files are accepted.  Matrices come out 0-based. */
struct acgb200_mtxinfo {
    int64_t nrows, ncols, nnzs;
    int64_t data_offset;      /* first byte after the size line */
    int field;                /* 0 real, 1 integer, 2 pattern, 3 other */
    int symmetric;
};
ACG_API int acgb200_mtx_info(const char *path, struct acgb200_mtxinfo *info);
ACG_API int acgb200_mtx_read(const char *path, int binary, struct acgsymcsrmatrix *A);
/* Part `part` of the row partition `rowparts` (nrows entries in [0,nparts)) of a
 * binary file, array for array what acgsymcsrmatrix_partition returns for that
 * part, without any process holding the whole matrix: the file is streamed and
 * only entries with an end in `part` are kept -- the per-rank replacement of the
 * reference's read-on-root + scatter (cuda/acg-cuda.c:1516-1782). */
ACG_API int acgb200_mtx_read_part(const char *path, int nparts, const int *rowparts, int part,
                                  struct acgsymcsrmatrix *A);

/* Row p of the communication matrix (cuda/acg-cuda.c:1713-1775, the driver's
 * --output-comm-matrix): row[q] = border values part A sends to part q per halo
 * exchange; nparts entries. */
ACG_API int acgb200_comm_matrix_row(const struct acgsymcsrmatrix *A, int nparts, int64_t *row);

/* Row-pattern dictionary of a 0-based CSR matrix (host logic of the pattern
 * slices, compress.c).  patptr needs max_entries+1 ints (at most that many
 * patterns), patoff max_entries ints, patid nrows entries; 0xFFFF in patid
 * marks a row whose pattern is not in the dictionary. */
ACG_API int acgb200_patterns_host(int nrows, const int64_t *rowptr, const int *colidx, int max_entries,
                                  int *npat, int *nentries, int *patptr, int *patoff, unsigned short *patid,
                                  int64_t *nmatched);

/* Inverse send map of the peer-memory halo exchange (host logic of p2p.c): for
 * border row b, entries [bptr[b], bptr[b+1]) give the recipient index bq[e] and
 * the destination offset bdst[e] in that recipient's ghost buffer;
 * rdispl_at_recipient[i] is recipient i's receive displacement for this sender.
 * bptr: nborder+1 ints, bq/bdst: halo->sendsize ints. */
ACG_API int acgb200_p2p_inverse_map(const struct acghalo *halo, int borderoff, int nborder,
                                    const int *rdispl_at_recipient, int *bptr, int *bq, int *bdst);

/* struct sizes, for bindings: "acgsolvercuda", "acgsymcsrmatrix", "acgvector",
 * "acgcomm", "acghalo"; returns 0 for unknown names */
ACG_API size_t acgb200_sizeof(const char *name);
/* 1 if this build was compiled with ACG_HAVE_MPI (struct acgcomm then has the
 * MPI member), else 0 */
ACG_API int acgb200_have_mpi(void);

/* page-lock / unlock caller-owned host memory (b->x, x->x) so that the solver's
 * H2D/D2H copies run at full PCIe speed; optional, the solve works on pageable memory */
ACG_API int acgb200_host_register(void *ptr, size_t bytes);
ACG_API int acgb200_host_unregister(void *ptr);

/* NCCL bootstrap for hosts without MPI (the reference driver broadcasts the
 * unique id over MPI, cuda/acg-cuda.c:1104-1122; bench.py / tests broadcast
 * these 128 bytes over torch.distributed instead) */
ACG_API int acgb200_nccl_unique_id(void *id128);
ACG_API int acgb200_comm_init_rank(struct acgcomm *comm, int nranks, const void *id128, int rank, int *ncclerrcode);
ACG_API int acgb200_comm_destroy(struct acgcomm *comm);

#ifdef __cplusplus
}
#endif
#endif
