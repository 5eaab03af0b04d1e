/*
 * internal.h -- private interfaces between the C host layer and the CUDA
 * kernels of libacgb200.  Not installed; the public C-ABI is include/acgb200/.
 */
#ifndef ACGB200_INTERNAL_H
#define ACGB200_INTERNAL_H

#include <cuda_runtime_api.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * SpMV tile plan
 *
 * The local CSR block is cut, once at init, into row-aligned tiles of at most
 * `rows_cap` rows and `nnz_cap` nonzeros.  A tile is one unit of TMA staging:
 * its slice of values / column indices / row pointers is fetched with three
 * cp.async.bulk copies into one shared-memory stage.  Bulk copies need
 * 16-byte aligned addresses and sizes, so every slice start is rounded down to
 * a multiple of 4 elements and every length rounded up; the device arrays are
 * allocated with padding so the over-read stays in bounds.
 *
 * Rows longer than nnz_cap ("long rows", power-law matrices) are left out of
 * the tiles and listed separately; a second kernel streams each of them with
 * several CTAs and combines partial sums with atomics.
 * ------------------------------------------------------------------------ */
struct acgb200_tile {        /* 16 bytes, read as one int4 */
    int row_begin;           /* first row of the tile */
    int nrows;               /* number of rows */
    int k_al;                /* rowptr[row_begin] rounded down to a multiple of 4 */
    int nnz_al;              /* padded slice length: multiple of 4, covers the tile's nonzeros */
};

struct acgb200_spmvplan {
    int nrows;                       /* rows covered (owned rows) */
    int64_t nnz;
    int lanes_per_row;               /* G: 1,2,4,8,16,32 */
    int rows_cap, nnz_cap;           /* tile limits */
    int nstages;                     /* smem ring depth */
    int threads;                     /* CTA size */
    int unroll;                      /* gathers in flight per lane */
    int ntiles;
    struct acgb200_tile *d_tiles;    /* [ntiles] */
    int nlong;                       /* rows with more than nnz_cap nonzeros */
    int *d_longrows;                 /* [nlong] */
    /* medium rows (opt-in, option "spmv_medium" = threshold): rows with more than med_thr and at most
     * nnz_cap nonzeros are left out of the tiles and get one warp each (spmv_medium_kernel); inside a
     * tile such a row would keep one G-lane group busy while the rest of the CTA idles */
    int med_thr;                     /* 0: off */
    int nmed;
    int *d_medrows;                  /* [nmed] */
    double *d_long_scratch;          /* [nlong * long_chunks] partial sums of the long-row kernels */
    int grid;                        /* persistent grid size */
    int smem_bytes;                  /* dynamic shared memory per CTA */
    int long_chunks;                 /* CTAs per long row */
    int max_ctas_per_sm;             /* 0: occupancy limit */
    /* pattern slices (slices.c): 32-row slices whose rows all repeat a dictionary pattern are stored
     * slice-major (entry e of the 32 rows contiguous, rows padded to the slice's longest) and multiplied by
     * spmv_slices_kernel -- 8 bytes per nonzero, no column indices, no row pointers, no shared-memory
     * staging; their rows are left out of the tiles */
    int nslices;
    struct acgb200_slice *d_slices;  /* [nslices] */
    double *d_sval;                  /* [sval_blocks * 32] */
    int64_t sval_blocks;             /* value blocks of 32 doubles (one per slice and entry slot) */
    int slice_rows;                  /* rows covered by slices */
    int64_t slice_nnz;               /* nonzeros covered (without padding) */
    int slice_lpad;                  /* row stride of the padded offset table */
    int slice_npat;
    int *d_spatoff;                  /* [slice_npat * slice_lpad] offsets col - row, zero beyond a pattern's length */
    unsigned short *d_spatid;        /* [nrows] pattern id per row (valid for covered rows) */
    int slice_ub, slice_threads, slice_grid, slice_smem, slice_max_ctas;
    int64_t slice_exc, slice_excnnz; /* exception rows inside slices (columns from the CSR index array) and their nonzeros */
    /* merge-path tiles (mergeplan.c): rows [0, merge_rows) of an irregular (power-law) matrix, cut into tiles
     * of merge_items merged items (row ends + nonzeros) and multiplied by spmv_merge_kernel; rows cut by tile
     * boundaries are finished by spmv_merge_fix_kernel from the tiles' partial sums */
    int nmtiles, nsplit, merge_items, merge_rows;
    int64_t merge_nnz;
    struct acgb200_mtile *d_mtiles;  /* [nmtiles] */
    struct acgb200_msplit *d_msplit; /* [nsplit] */
    double *d_mpart;                 /* [2 * nmtiles]: slot 0 head piece, slot 1 tail piece of each tile */
    int merge_grid, merge_smem, merge_threads, merge_stages, merge_max_ctas;
};

/* one merge-path tile: row ends of rows [r0, r0+nre), nonzeros [k0, k0+nnz) */
struct acgb200_mtile { int r0; int nre; int k0; int nnz; };
/* a row cut by tile boundaries: its pieces are the tail pieces of tiles ta..tb-1 and the head piece of tile tb */
struct acgb200_msplit { int row; int ta; int tb; int pad; };
struct acgb200_mergeplan {
    int ntiles, nsplit, items, rows;
    int64_t nnz;
    struct acgb200_mtile *tiles;
    struct acgb200_msplit *split;
};
int acgb200_merge_plan(int hi, const int64_t *rowptr, int items, struct acgb200_mergeplan *out);
void acgb200_mergeplan_free(struct acgb200_mergeplan *mp);

struct acgb200_patterns;

/* one covered slice: rows [row0, row0+32), `len` entry slots, values at d_sval + 32 * vblk */
struct acgb200_slice { int row0; int nrows; int len; int vblk; };

/* host plan of the pattern slices (no CUDA): which full 32-row slices of rows [0,cover_hi) are
 * covered, their descriptors, and the padded offset table.  Returns ACG_SUCCESS with *nslices == 0
 * when slices do not pay (few covered rows).  Arrays are malloc'ed, the caller frees them. */
struct acgb200_sliceplan {
    int nslices; struct acgb200_slice *slices;
    unsigned char *covered;          /* [(nrows+31)/32] 1: the slice's rows are handled by the slice kernel */
    int64_t blocks, nnz;
    int rows, lpad, npat, domlen;    /* domlen: most frequent row length among covered rows */
    int *spatoff;                    /* [npat * lpad] */
    unsigned short *patid;           /* [nrows] pattern id as the slice kernel sees it: < npat, or ACGB200_NOPATTERN =
                                      * exception row (columns from the index array) */
    int64_t nexc, excnnz;            /* exception rows inside covered slices, their nonzeros */
};
int acgb200_slices_plan(int nrows, int cover_hi, const int64_t *rowptr, const struct acgb200_patterns *pat,
                        struct acgb200_sliceplan *out);
void acgb200_sliceplan_free(struct acgb200_sliceplan *sp);
/* fill d_sval from the device CSR values (kernels.cu) */
int acgb200_slices_fill(const struct acgb200_spmvplan *pl, const int *d_rowptr, const double *d_a, cudaStream_t stream);

/* row-pattern dictionary (compress.c): rows whose pattern id is not
 * ACGB200_NOPATTERN need no column indices, col = row + patoff[patptr[id] + j] */
#define ACGB200_NOPATTERN 0xFFFFu
struct acgb200_patterns {
    int npat, nentries;
    int *patptr;                 /* [npat+1] */
    int *patoff;                 /* [nentries] */
    unsigned short *patid;       /* [nrows] */
    int64_t nrows_matched;
};
int acgb200_patterns_build(int nrows, const int64_t *rowptr, const int *colidx, int max_entries,
                           struct acgb200_patterns *out);
void acgb200_patterns_free(struct acgb200_patterns *p);

/* epilogue of the SpMV kernels */
enum acgb200_spmvmode {
    SPMV_Y_AX = 0,        /* y = A x                                            */
    SPMV_Y_AX_DOT = 1,    /* y = A x ; acc += sum_{row<dotrows} x[row]*y[row]   */
    SPMV_R_B_AX = 2,      /* y = b - A x ; acc += sum_{row<dotrows} y[row]^2    */
    SPMV_Y_PLUS_AX = 3,   /* y += A x  (off-diagonal block)                     */
};

/* ------------------------------------------------------------------------
 * Device-resident CG state: iteration control and all scalars.  One
 * allocation; kernels never read a control word that the same kernel writes
 * (see DESIGN.md "control ring").
 * ------------------------------------------------------------------------ */
struct acgb200_ctrl { int iter; int done; int pad0, pad1; };

struct acgb200_devstate {
    struct acgb200_ctrl ctrl[4];     /* ring: kernel i of an iteration reads ctrl[i], writes ctrl[i+1 mod nk] */
    int maxits;
    int pad[3];
    double tol;                      /* stop when ||r|| < tol (0: never) */
    double final_rr;                 /* (r,r) at the iteration that set done */
    /* classic CG: slot = iteration parity */
    double pap_loc[2], pap[2];       /* local partial / global (p,Ap) */
    double rr_loc[2], rr[2];         /* (r_k,r_k) lives in slot k&1 */
    /* pipelined CG */
    double gd_loc[2][2], gd[2][2];   /* {gamma,delta} = {(r,r),(w,r)} */
    double prev[2][2];               /* {gamma_{k-1}, alpha_{k-1}} read by update k from slot k&1 */
    /* setup-time reductions */
    double tmp_loc[2], tmp[2];
    unsigned int ticket;             /* spare (last-CTA detection on one GPU) */
    unsigned int pad1;
};

/* ------------------------------------------------------------------------
 * Peer-memory exchange (NVLink 5 / NVSwitch, one process per GPU).
 *
 * Every rank exports one device allocation (the "window") through CUDA IPC:
 *
 *     uint64 hflag[MAXR]                halo sequence number, per sender rank
 *     uint64 rflag[NCH][MAXR]           reduction sequence number, per channel/rank
 *     double red[NCH][2][MAXR][2]       reduction partials [channel][parity][rank][2]
 *     double ghost[2][recvsize]         ghost values, double-buffered by parity
 *
 * Producers store straight into their peers' windows (remote stores over
 * NVLink) and then publish a sequence number with st.release.sys; consumers
 * spin with ld.acquire.sys on their own window.  This replaces, inside the CG
 * loop, the pack kernel + ncclSend/ncclRecv group + unpack of the halo exchange
 * (acg/halo.c:1456-1627) and the ncclAllReduce of 1-2 doubles
 * (acg/comm.c:371): see DESIGN.md section 6.
 * ------------------------------------------------------------------------ */
#define ACGB200_MAXR 64
#define ACGB200_NCH 2

struct acgb200_p2pdev {
    int nranks, rank;
    int nrecip, sendsize;                    /* halo send side: neighbours, entries */
    int sdispls[ACGB200_MAXR + 1];           /* my send segments */
    int peer_rdispl[ACGB200_MAXR];           /* where segment i starts in recipient i's ghost buffer */
    double *peer_ghost[ACGB200_MAXR][2];     /* recipient i's ghost buffers */
    unsigned long long *peer_hflag[ACGB200_MAXR];   /* recipient i's hflag slot for this rank */
    int nsenders;
    int senders[ACGB200_MAXR];               /* ranks this rank waits for */
    unsigned long long *my_hflag;            /* [MAXR] */
    double *my_ghost[2];
    double *peer_red[ACGB200_MAXR];          /* rank r's red[][][][] base */
    unsigned long long *peer_rflag[ACGB200_MAXR];   /* rank r's rflag[][] base */
    double *my_red;
    unsigned long long *my_rflag;
    unsigned long long hbase, rbase;         /* sequence bases of the current solve */
    unsigned long long timed_out;            /* set by a kernel whose wait for a peer exceeded timeout_ns; sticky
                                              * for the solve (hbase, rbase, timed_out are reset by one copy) */
    unsigned long long timeout_ns;           /* 0: wait for ever */
    unsigned int ticket;                     /* last-block detection (one kernel at a time uses it) */
    /* inverse send map: border row b (relative to borderrowoffset) is sent to
     * neighbours bq[e] at ghost offsets bdst[e], e in [bptr[b], bptr[b+1]) */
    int borderoff, nborder;
    const int *bptr, *bq, *bdst;
    int fuse;                                /* 1: producers push/publish themselves; 0: comm_post_kernel does */
};

/* After the producer of a vector / of reduction partials: push them to the
 * peers.  iter_override < 0: gated by *cin and using its iteration index. */
struct acgb200_postargs {
    struct acgb200_p2pdev *p2p;
    const struct acgb200_ctrl *cin;
    const struct acgb200_devstate *st;
    int iter_override;
    const double *vec; const int *sendbufidx;     /* halo part; vec == NULL: none */
    int ch; const double *redbase; int redstride, redcount, par_off, seq_off;   /* reduction part; ch < 0: none */
};
int acgb200_comm_post(const struct acgb200_postargs *a, cudaStream_t stream);

/* kernels.cu ------------------------------------------------------------- */

/* choose tile parameters for a matrix with the given shape; fills everything
 * in *plan except the device arrays */
void acgb200_spmv_choose(struct acgb200_spmvplan *plan, int nrows, int64_t nnz, int64_t maxrowlen);
int acgb200_spmv_configure(struct acgb200_spmvplan *plan);   /* occupancy -> grid; returns cudaError */

struct acgb200_spmvargs {
    const struct acgb200_spmvplan *plan;
    const int *rowptr;        /* [nrows+1] (+pad) */
    const int *colidx;        /* [nnz] (+pad) */
    const double *a;          /* [nnz] (+pad) */
    const double *x;
    double *y;
    const double *b;          /* SPMV_R_B_AX */
    double *acc;              /* device accumulator for the fused dot */
    int dotrows;              /* rows [0,dotrows) contribute to acc */
    int mode;
    const struct acgb200_ctrl *ctrl_in;   /* may be NULL: unconditional */
    struct acgb200_ctrl *ctrl_out;
    struct acgb200_devstate *st;          /* for per-iteration scalar housekeeping; may be NULL */
    int housekeeping;                     /* 0 none, 1 classic, 2 pipelined */
    /* peer-memory mode: the border x ghost block is applied inside the kernel
     * (ghost values from the window once the senders have published) and the
     * last CTA publishes the fused dot on reduction channel pub_ch */
    const struct acgb200_p2pdev *p2p;
    int od_rowoffset, od_nrows;
    const int *orowptr; const int *ocolidx; const double *oa;
    int pub_ch;                           /* -1: do not publish */
};
/* bytes one SpMV launch must move at least, given the plan (for reports) */
int64_t acgb200_spmv_min_bytes(const struct acgb200_spmvplan *plan);
int acgb200_spmv_launch(const struct acgb200_spmvargs *args, cudaStream_t stream);

/* y[rowoffset+i] (+)= sum_k oa[k]*x[xoffset + ocolidx[k]], i in [0,nrows): the
 * border x ghost block (acg/cgcuda.c:878); epilogue per `mode` */
struct acgb200_offdiagargs {
    int nrows;                /* border rows */
    int rowoffset;            /* borderrowoffset */
    const int *orowptr; const int *ocolidx; const double *oa;
    const double *x; double *y;
    double *acc; int mode;    /* SPMV_Y_PLUS_AX (+ dot with x over these rows if acc) or SPMV_R_B_AX-style minus */
    int minus;                /* y -= instead of += (residual) */
    int dotkind;              /* 0 none, 1 x[row]*y[row], 2 y[row]^2 */
    const struct acgb200_ctrl *ctrl_in;
    struct acgb200_ctrl *ctrl_out;
    struct acgb200_devstate *st;
    const struct acgb200_p2pdev *p2p;   /* not NULL: ghosts come from the peer-memory window */
    int p2p_iter_override;              /* >= 0: wait for that halo sequence index (setup) */
};
int acgb200_offdiag_launch(const struct acgb200_offdiagargs *args, cudaStream_t stream);

/* classic CG fused BLAS-1 (replace acg/cg-kernels-cuda.cu:119-303 and the two
 * cublasDdot calls at acg/cgcuda.c:894,933) */
/* p2p != NULL: reductions come from / go to the peer-memory window and the
 * kernels that produce the next SpMV input push its border entries themselves */
int acgb200_cg_update_r(int n, struct acgb200_devstate *st, int cin, int cout, int multi,
                        struct acgb200_p2pdev *p2p,
                        const double *t, double *r, cudaStream_t stream);
int acgb200_cg_update_xp(int n, struct acgb200_devstate *st, int cin, int cout, int multi,
                         struct acgb200_p2pdev *p2p,
                         const double *r, double *p, double *x, cudaStream_t stream);
/* pipelined CG fused update + next dots (replaces acg/cg-kernels-cuda.cu:187-269
 * and the cublasDdot calls at acg/cgcuda.c:1680,1688) */
int acgb200_pcg_update(int n, struct acgb200_devstate *st, int cin, int cout, int multi,
                       struct acgb200_p2pdev *p2p,
                       const double *q, double *z, double *w, double *t, double *p,
                       double *r, double *x, cudaStream_t stream);

/* the reference's public BLAS-1 building blocks (acg/cg-kernels-cuda.h:45-97; cgcuda.c wraps them) */
int acgb200_helper_axpy(int op, int n, const double *num, const double *den, const double *x, double *y, cudaStream_t stream);
int acgb200_helper_scalars(int op, double *out0, double *out1, const double *num, const double *den, cudaStream_t stream);
int acgb200_helper_pipelined(int n, const double *gamma, double *gamma_prev, const double *delta, const double *q,
                             double *p, double *r, double *t, double *x, double *z, double *w, double *alpha_prev,
                             cudaStream_t stream);

/* expand.cu: full-storage expansion of the packed symmetric CSR on the device
 * (the GPU form of acgsymcsrmatrix_dsymv_init, acg/symcsrmatrix.c:760-851) */
struct acgb200_expanded {
    int *d_rowptr; int *d_colidx; double *d_a;          /* local block, 0-based columns */
    int *d_orowptr; int *d_ocolidx; double *d_oa;       /* border x ghost block, columns rebased by -borderrowoffset */
    int64_t fnnz, onnz;
};
int acgb200_expand_device(int n, int ghost0, int border0, int nob, int64_t pnnz,
                          const int *d_prp, const int *d_pcol, const double *d_pa, double eps,
                          int rp_pad, int blk_pad, struct acgb200_expanded *out, cudaStream_t stream);

struct acgsymcsrmatrix;
/* expand_host.c: upload the packed matrix and expand it on the device (rp_pad / blk_pad: trailing
 * padding of the row-pointer / column+value arrays, as the TMA tile copies need) */
int acgb200_expand_upload(const struct acgsymcsrmatrix *A, double eps, int rp_pad, int blk_pad,
                          struct acgb200_expanded *out, cudaStream_t stream, int *errcode);
void acgb200_expanded_free(struct acgb200_expanded *x);

/* setup-time helpers */
int acgb200_dot(int n, const double *x, const double *y, double *acc, cudaStream_t stream); /* acc += x.y */
int acgb200_dot2(int n, const double *r, const double *w, double *acc2, cudaStream_t stream); /* acc2[0]+=r.r, acc2[1]+=w.r */
int acgb200_gather(int n, double *dst, const double *src, const int *idx, cudaStream_t stream);  /* dst[i]=src[idx[i]] */
int acgb200_scatter(int n, const double *src, double *dst, const int *idx, cudaStream_t stream); /* dst[idx[i]]=src[i] */
int acgb200_num_sms(void);
void acgb200_blas1_set_ctas_per_sm(int v);
void acgb200_set_pdl(int v);

#ifdef __cplusplus
}
#endif
#endif
