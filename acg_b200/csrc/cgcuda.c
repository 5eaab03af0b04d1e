/*
 * cgcuda.c -- host side of the B200 conjugate-gradient solver (C).
 *
 * Implements the drop-in boundary declared in include/acgb200/cgcuda.h:
 * acgsolvercuda_init / _solvempi / _solve_pipelined / _solve / _fwrite /
 * _free (reference: acg/cgcuda.c).  The host only sequences kernels and
 * NCCL calls; all arithmetic happens in kernels.cu.  There is no CPU
 * fallback: every entry point fails with ACG_ERR_CUDA if no device is usable.
 *
 * Per-iteration schedule, classic CG (replaces acg/cgcuda.c:845-1019):
 *
 *   comm stream : [pack border of p -> ncclSend/Recv into ghost tail of p]
 *   main stream : t = A_loc p (+ p.t over interior rows)
 *                 [wait halo; t += A_off p on border rows (+ p.t over border)]
 *                 [allreduce p.t]
 *                 r -= alpha t  (+ r.r)           alpha = r.r_old / p.t
 *                 [allreduce r.r]
 *                 x += alpha p ; p = r + beta p   beta = r.r / r.r_old
 *
 * Scalars never leave the device.  Convergence is decided on the device by
 * the last kernel of each iteration, which also advances an iteration
 * counter; once "done" is set every later kernel returns immediately, so the
 * host may run ahead and only polls the control word every few iterations
 * (the reference synchronises the host on every iteration, :1007).
 */
#include "acgb200/cgcuda.h"
#include "acgb200/error.h"
#include "acgb200/ext.h"
#include "internal.h"
#include "p2p.h"

#include <cuda_runtime_api.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------ */
/* tunables                                                                  */
/* ------------------------------------------------------------------------ */

#define SPMV_MAX_STAGES_HOST 8      /* SPMV_MAX_STAGES of kernels.cu */

static struct {
    int profile;        /* record CUDA events around each kernel class */
    int check_every;    /* iterations between convergence polls */
    int spmv_lanes;     /* 0 = heuristic */
    int spmv_nnz_cap, spmv_rows_cap, spmv_stages, spmv_threads, spmv_unroll;
    int spmv_max_ctas;  /* cap on resident SpMV CTAs per SM (0 = occupancy limit) */
    int graph;          /* replay iteration pairs as CUDA graphs */
    int redstream;      /* pipelined CG: allreduce on its own stream + communicator */
    int p2p;            /* halo + reductions through peer memory (CUDA IPC) instead of NCCL */
    int p2p_fuse;       /* 1: border x ghost block inside the SpMV, pushes inside the update kernels */
    int blas1_ctas;     /* CTAs per SM of the fused BLAS-1 kernels (0 = one full wave, from the occupancy) */
    int pdl;            /* 1: programmatic dependent launch along the iteration chain (opt-in) */
    int spmv_medium;    /* > 0: rows longer than this (and shorter than a tile) get a warp each (opt-in) */
    int spmv_merge;     /* merge-path tiles for irregular rows (mergeplan.c): -1 decide from the row lengths, 0 off, 1 on */
    int merge_items, merge_threads, merge_max_ctas, merge_stages;   /* their shape (0 = default) */
    int spmv_slices;    /* 1: pattern slices (slices.c) -- index-free slice-major storage of the rows that repeat a pattern */
    int slice_ub, slice_threads, slice_max_ctas;   /* slice kernel shape overrides (0 = default) */
    int loaded;
} cfg = { .check_every = 8, .graph = 1, .redstream = 1, .p2p = 1, .p2p_fuse = 1, .spmv_slices = 1, .spmv_merge = -1 };

static void cfg_load(void)
{
    if (cfg.loaded) return;
    cfg.loaded = 1;
    const char *s;
    if ((s = getenv("ACGB200_PROFILE"))) cfg.profile = atoi(s);
    if ((s = getenv("ACGB200_CHECK_EVERY"))) cfg.check_every = atoi(s);
    if ((s = getenv("ACGB200_SPMV_LANES"))) cfg.spmv_lanes = atoi(s);
    if ((s = getenv("ACGB200_SPMV_NNZ_CAP"))) cfg.spmv_nnz_cap = atoi(s);
    if ((s = getenv("ACGB200_SPMV_ROWS_CAP"))) cfg.spmv_rows_cap = atoi(s);
    if ((s = getenv("ACGB200_SPMV_STAGES"))) cfg.spmv_stages = atoi(s);
    if ((s = getenv("ACGB200_SPMV_THREADS"))) cfg.spmv_threads = atoi(s);
    if ((s = getenv("ACGB200_SPMV_UNROLL"))) cfg.spmv_unroll = atoi(s);
    if ((s = getenv("ACGB200_SPMV_MAX_CTAS"))) cfg.spmv_max_ctas = atoi(s);
    if ((s = getenv("ACGB200_GRAPH"))) cfg.graph = atoi(s);
    if ((s = getenv("ACGB200_REDSTREAM"))) cfg.redstream = atoi(s);
    if ((s = getenv("ACGB200_P2P"))) cfg.p2p = atoi(s);
    if ((s = getenv("ACGB200_P2P_FUSE"))) cfg.p2p_fuse = atoi(s);
    if ((s = getenv("ACGB200_BLAS1_CTAS"))) cfg.blas1_ctas = atoi(s);
    if (cfg.check_every < 1) cfg.check_every = 1;
    if ((s = getenv("ACGB200_PDL"))) cfg.pdl = atoi(s);
    if ((s = getenv("ACGB200_SPMV_MEDIUM"))) cfg.spmv_medium = atoi(s);
    if ((s = getenv("ACGB200_SPMV_MERGE"))) cfg.spmv_merge = atoi(s);
    if ((s = getenv("ACGB200_MERGE_ITEMS"))) cfg.merge_items = atoi(s);
    if ((s = getenv("ACGB200_MERGE_THREADS"))) cfg.merge_threads = atoi(s);
    if ((s = getenv("ACGB200_MERGE_MAX_CTAS"))) cfg.merge_max_ctas = atoi(s);
    if ((s = getenv("ACGB200_MERGE_STAGES"))) cfg.merge_stages = atoi(s);
    if ((s = getenv("ACGB200_SPMV_SLICES"))) cfg.spmv_slices = atoi(s);
    if ((s = getenv("ACGB200_SLICE_UB"))) cfg.slice_ub = atoi(s);
    if ((s = getenv("ACGB200_SLICE_THREADS"))) cfg.slice_threads = atoi(s);
    if ((s = getenv("ACGB200_SLICE_MAX_CTAS"))) cfg.slice_max_ctas = atoi(s);
    acgb200_blas1_set_ctas_per_sm(cfg.blas1_ctas);
    acgb200_set_pdl(cfg.pdl);
}

int acgb200_set_option(const char *key, int value)
{
    cfg_load();
    if (!strcmp(key, "profile")) cfg.profile = value;
    else if (!strcmp(key, "check_every")) cfg.check_every = value < 1 ? 1 : value;
    else if (!strcmp(key, "spmv_lanes")) cfg.spmv_lanes = value;
    else if (!strcmp(key, "spmv_nnz_cap")) cfg.spmv_nnz_cap = value;
    else if (!strcmp(key, "spmv_rows_cap")) cfg.spmv_rows_cap = value;
    else if (!strcmp(key, "spmv_stages")) cfg.spmv_stages = value;
    else if (!strcmp(key, "spmv_threads")) cfg.spmv_threads = value;
    else if (!strcmp(key, "spmv_unroll")) cfg.spmv_unroll = value;
    else if (!strcmp(key, "spmv_max_ctas")) cfg.spmv_max_ctas = value;
    else if (!strcmp(key, "graph")) cfg.graph = value;
    else if (!strcmp(key, "redstream")) cfg.redstream = value;
    else if (!strcmp(key, "p2p")) cfg.p2p = value;
    else if (!strcmp(key, "p2p_fuse")) cfg.p2p_fuse = value;
    else if (!strcmp(key, "blas1_ctas")) { cfg.blas1_ctas = value; acgb200_blas1_set_ctas_per_sm(value); }
    else if (!strcmp(key, "pdl")) { cfg.pdl = value; acgb200_set_pdl(value); }
    else if (!strcmp(key, "spmv_medium")) cfg.spmv_medium = value < 0 ? 0 : value;
    else if (!strcmp(key, "spmv_merge")) cfg.spmv_merge = value;
    else if (!strcmp(key, "merge_items")) cfg.merge_items = value;
    else if (!strcmp(key, "merge_threads")) cfg.merge_threads = value;
    else if (!strcmp(key, "merge_max_ctas")) cfg.merge_max_ctas = value;
    else if (!strcmp(key, "merge_stages")) cfg.merge_stages = value;
    else if (!strcmp(key, "spmv_slices")) cfg.spmv_slices = value;
    else if (!strcmp(key, "slice_ub")) cfg.slice_ub = value;
    else if (!strcmp(key, "slice_threads")) cfg.slice_threads = value;
    else if (!strcmp(key, "slice_max_ctas")) cfg.slice_max_ctas = value;
    else return ACG_ERR_INVALID_VALUE;
    return ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* private per-solver state                                                  */
/* ------------------------------------------------------------------------ */

struct evpool { cudaEvent_t *ev; int n, cap; };

struct priv {
    const struct acgsolvercuda *key;
    const void *d_r_id;                 /* cg->d_r: identifies the solver if the caller moves the struct */
    struct priv *next;
    struct acgb200_spmvplan plan;
    struct acgb200_devstate *d_st;
    struct acgb200_ctrl *h_ctrl;        /* pinned, 2 poll slots */
    struct acgb200_devstate *h_st;      /* pinned scratch for state upload / readback */
    int nowned, ninner, nborder, nghost, borderoff, nvec;
    int64_t fnnz, onnz;
    int device_expanded;                /* the full storage was built on the device (the host matrix has none) */
    cudaStream_t stream, commstream, redstream;
    cudaEvent_t ev_ready, ev_halo, ev_red, ev_poll[2];
    struct acgb200_p2p p2p;             /* peer-memory exchange (multi-GPU) */
    struct acgcomm redcomm;             /* private duplicate of the caller's communicator for reductions */
    int have_redcomm;
    double *d_b, *d_x;                  /* right-hand side / solution on the device, kept between solves */
    cudaGraphExec_t graph[3];           /* [0] classic, [1] pipelined ([2] spare): two iterations each (parity 0 then 1) */
    int graph_sig[3];                   /* loop configuration each cached graph was captured with (graph_signature) */
    int graph_launches[3];              /* kernel/NCCL launches inside one replay */
    double last_h2d_ms, last_d2h_ms;    /* host time spent before / after the solve window */
    cudaEvent_t ev_t0, ev_t1;           /* device-side bracket of the solve window */
    double last_solve_ms;
    struct evpool gemv, blas;           /* profiling */
    int last_launches;                  /* kernels launched in the last solve's timed loop */
    double last_spmv_ms;                /* profiled SpMV time of the last solve */
    double last_blas_ms;                /* profiled fused-update time of the last solve */
    int last_spmv_n;
};

/* The public struct has no room for private state (its layout is the
 * reference's), so it lives in a registry keyed by the solver's address.  The
 * lock only guards the list: one solver is driven by one host thread, as in the
 * reference, but different solvers may live on different threads. */
static struct priv *registry = NULL;
static pthread_mutex_t registry_lock = PTHREAD_MUTEX_INITIALIZER;

static struct priv *priv_of(const struct acgsolvercuda *cg)
{
    struct priv *found = NULL;
    pthread_mutex_lock(&registry_lock);
    for (struct priv *p = registry; p; p = p->next) if (p->key == cg) { found = p; break; }
    if (!found && cg->d_r) {
        /* the caller copied or moved the struct (the reference allows that: it is plain data).  The device
         * vector d_r is allocated by acgsolvercuda_init and belongs to exactly one solver: find the state
         * through it and follow the struct to its new address. */
        for (struct priv *p = registry; p; p = p->next)
            if (p->d_r_id == (const void *) cg->d_r) { found = p; p->key = cg; break; }
    }
    pthread_mutex_unlock(&registry_lock);
    return found;
}

static void priv_add(struct priv *pv)
{
    pthread_mutex_lock(&registry_lock);
    pv->next = registry; registry = pv;
    pthread_mutex_unlock(&registry_lock);
}

static void priv_drop(struct priv *pv)
{
    struct priv *d = NULL;
    pthread_mutex_lock(&registry_lock);
    for (struct priv **pp = &registry; *pp; pp = &(*pp)->next) {
        if (*pp == pv) { d = *pp; *pp = d->next; break; }
    }
    pthread_mutex_unlock(&registry_lock);
    free(d);
}

static double wall(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (errcode) *errcode = (int) e_; return ACG_ERR_CUDA; } } while (0)
#define KL(call) do { int e_ = (call); if (e_) { if (errcode) *errcode = e_; return ACG_ERR_CUDA; } } while (0)
#define OK(call) do { int e_ = (call); if (e_) return e_; } while (0)

/* ------------------------------------------------------------------------ */
/* free / init                                                               */
/* ------------------------------------------------------------------------ */

static void free_vecptr(struct acgvector **v)
{
    if (*v) { acgvector_free(*v); free(*v); *v = NULL; }
}

void acgsolvercuda_free(struct acgsolvercuda *cg)
{
    struct priv *pv = priv_of(cg);
    acgvector_free(&cg->r); acgvector_free(&cg->p); acgvector_free(&cg->t);
    free_vecptr(&cg->w); free_vecptr(&cg->q); free_vecptr(&cg->z); free_vecptr(&cg->dx);
    if (cg->halo) { acghalo_free(cg->halo); free(cg->halo); cg->halo = NULL; }
    if (cg->haloexchange) { acghaloexchange_free(cg->haloexchange); free(cg->haloexchange); cg->haloexchange = NULL; }
    cudaFree(cg->d_r); cudaFree(cg->d_p); cudaFree(cg->d_t);
    cudaFree(cg->d_w); cudaFree(cg->d_q); cudaFree(cg->d_z);
    cudaFree(cg->d_rowptr); cudaFree(cg->d_colidx); cudaFree(cg->d_a);
    cudaFree(cg->d_orowptr); cudaFree(cg->d_ocolidx); cudaFree(cg->d_oa);
    cg->d_r = cg->d_p = cg->d_t = cg->d_w = cg->d_q = cg->d_z = NULL;
    cg->d_rowptr = cg->d_colidx = cg->d_orowptr = cg->d_ocolidx = NULL;
    cg->d_a = cg->d_oa = NULL;
    if (pv) {
        if (pv->stream) cudaStreamSynchronize(pv->stream);
        if (pv->p2p.window) {
            /* unmap the peers' windows, wait until every rank has done so, then
             * release this rank's (collective, like the init that created it) */
            acgb200_p2p_free(&pv->p2p);
            if (pv->have_redcomm) { acgcomm_barrier(pv->stream, &pv->redcomm, NULL); cudaStreamSynchronize(pv->stream); }
            cudaFree(pv->p2p.window);
            pv->p2p.window = NULL;
        }
        if (pv->have_redcomm && pv->redcomm.ncclcomm) ncclCommDestroy(pv->redcomm.ncclcomm);
        for (int i = 0; i < 3; i++) if (pv->graph[i]) cudaGraphExecDestroy(pv->graph[i]);
        cudaFree(pv->plan.d_tiles); cudaFree(pv->plan.d_longrows); cudaFree(pv->plan.d_long_scratch); cudaFree(pv->plan.d_medrows);
        cudaFree(pv->plan.d_slices); cudaFree(pv->plan.d_sval); cudaFree(pv->plan.d_spatoff); cudaFree(pv->plan.d_spatid);
        cudaFree(pv->plan.d_mtiles); cudaFree(pv->plan.d_msplit); cudaFree(pv->plan.d_mpart);
        cudaFree(pv->d_st);
        cudaFreeHost(pv->h_ctrl); cudaFreeHost(pv->h_st);
        if (pv->stream) cudaStreamDestroy(pv->stream);
        if (pv->commstream) cudaStreamDestroy(pv->commstream);
        if (pv->redstream) cudaStreamDestroy(pv->redstream);
        if (pv->ev_red) cudaEventDestroy(pv->ev_red);
        cudaFree(pv->d_b); cudaFree(pv->d_x);
        if (pv->ev_ready) cudaEventDestroy(pv->ev_ready);
        if (pv->ev_halo) cudaEventDestroy(pv->ev_halo);
        for (int i = 0; i < 2; i++) if (pv->ev_poll[i]) cudaEventDestroy(pv->ev_poll[i]);
        if (pv->ev_t0) cudaEventDestroy(pv->ev_t0);
        if (pv->ev_t1) cudaEventDestroy(pv->ev_t1);
        for (int i = 0; i < pv->gemv.cap; i++) cudaEventDestroy(pv->gemv.ev[i]);
        for (int i = 0; i < pv->blas.cap; i++) cudaEventDestroy(pv->blas.ev[i]);
        free(pv->gemv.ev); free(pv->blas.ev);
        priv_drop(pv);
    }
}

/* Cut rows [0,nrows) into TMA tiles (see internal.h): greedy, row-aligned, at
 * most rows_cap rows and nnz_cap nonzeros per tile; rows longer than nnz_cap go
 * to the long-row list.  Host-only, no CUDA: testable without a device. */
static int cut_tiles(const struct acgb200_spmvplan *pl, const int64_t *rowptr, const unsigned char *covered, int row_lo,
                     struct acgb200_tile *tiles, int *ntiles, int *longrows, int *nlong, int *medrows, int *nmed)
{
    const int n = pl->nrows;
    /* rows above `out` leave the tiles: the long ones (> nnz_cap) always, the medium ones on request */
    const int64_t out = pl->med_thr > 0 && pl->med_thr < pl->nnz_cap ? pl->med_thr : pl->nnz_cap;
    int nt = 0, nl = 0, nm = 0, r = row_lo;        /* rows below row_lo are in merge-path tiles (mergeplan.c) */
    while (r < n) {
        /* rows of a covered 32-row slice belong to the slice kernel (slices.c) */
        if (covered && covered[r >> 5]) { r = (r | 31) + 1; continue; }
        int64_t len = rowptr[r + 1] - rowptr[r];
        if (len > pl->nnz_cap) { longrows[nl++] = r++; continue; }
        if (len > out) { medrows[nm++] = r++; continue; }
        const int start = r;
        int64_t cnt = 0;
        while (r < n && r - start < pl->rows_cap) {
            if (covered && covered[r >> 5]) break;
            len = rowptr[r + 1] - rowptr[r];
            if (len > out || cnt + len > pl->nnz_cap) break;
            cnt += len; r++;
        }
        const int64_t kb = rowptr[start], ke = rowptr[r];
        if (ke > INT32_MAX) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
        const int64_t k_al = kb & ~(int64_t) 3;
        tiles[nt].row_begin = start;
        tiles[nt].nrows = r - start;
        tiles[nt].k_al = (int) k_al;
        tiles[nt].nnz_al = (int) (((ke - k_al) + 3) & ~(int64_t) 3);
        nt++;
    }
    *ntiles = nt; *nlong = nl; *nmed = nm;
    return ACG_SUCCESS;
}

static int build_tiles(struct acgb200_spmvplan *pl, const int64_t *rowptr,
                       const struct acgb200_sliceplan *sp, const unsigned short *patid,
                       const struct acgb200_mergeplan *mp, int *errcode)
{
    const int n = pl->nrows;
    struct acgb200_tile *tiles = malloc(((size_t) n + 1) * sizeof(*tiles));
    int *longrows = malloc(((size_t) n + 1) * sizeof(*longrows));
    int *medrows = malloc(((size_t) n + 1) * sizeof(*medrows));
    if (!tiles || !longrows || !medrows) { free(tiles); free(longrows); free(medrows); return ACG_ERR_ERRNO; }
    int nt = 0, nl = 0, nm = 0;
    int err = cut_tiles(pl, rowptr, sp && sp->nslices > 0 ? sp->covered : NULL, mp && mp->ntiles > 0 ? mp->rows : 0,
                        tiles, &nt, longrows, &nl, medrows, &nm);
    if (err) { free(tiles); free(longrows); free(medrows); return err; }
    pl->ntiles = nt; pl->nlong = nl; pl->nmed = nm;
    pl->d_tiles = NULL; pl->d_longrows = NULL; pl->d_long_scratch = NULL; pl->d_medrows = NULL;
    cudaError_t e = cudaSuccess;
    if (!e && nt > 0) {
        e = cudaMalloc((void **) &pl->d_tiles, (size_t) nt * sizeof(*tiles));
        if (!e) e = cudaMemcpy(pl->d_tiles, tiles, (size_t) nt * sizeof(*tiles), cudaMemcpyHostToDevice);
    }
    if (!e && nl > 0) {
        e = cudaMalloc((void **) &pl->d_longrows, (size_t) nl * sizeof(int));
        if (!e) e = cudaMemcpy(pl->d_longrows, longrows, (size_t) nl * sizeof(int), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_long_scratch, (size_t) nl * (size_t) pl->long_chunks * sizeof(double));
    }
    if (!e && nm > 0) {
        e = cudaMalloc((void **) &pl->d_medrows, (size_t) nm * sizeof(int));
        if (!e) e = cudaMemcpy(pl->d_medrows, medrows, (size_t) nm * sizeof(int), cudaMemcpyHostToDevice);
    }
    pl->nslices = 0;
    if (!e && sp && sp->nslices > 0) {
        pl->nslices = sp->nslices; pl->sval_blocks = sp->blocks; pl->slice_rows = sp->rows; pl->slice_nnz = sp->nnz;
        pl->slice_lpad = sp->lpad; pl->slice_npat = sp->npat; pl->slice_exc = sp->nexc; pl->slice_excnnz = sp->excnnz;
        patid = sp->patid;          /* ids as the slice kernel sees them (exception rows marked) */
        const size_t tab = (size_t) sp->npat * (size_t) sp->lpad;
        e = cudaMalloc((void **) &pl->d_slices, (size_t) sp->nslices * sizeof(*sp->slices));
        if (!e) e = cudaMemcpy(pl->d_slices, sp->slices, (size_t) sp->nslices * sizeof(*sp->slices), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_spatoff, (tab + 4) * sizeof(int));
        if (!e) e = cudaMemset(pl->d_spatoff, 0, (tab + 4) * sizeof(int));
        if (!e) e = cudaMemcpy(pl->d_spatoff, sp->spatoff, tab * sizeof(int), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_spatid, ((size_t) n + 32) * sizeof(unsigned short));
        if (!e) e = cudaMemset(pl->d_spatid, 0, ((size_t) n + 32) * sizeof(unsigned short));
        if (!e) e = cudaMemcpy(pl->d_spatid, patid, (size_t) n * sizeof(unsigned short), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_sval, (size_t) sp->blocks * 32 * sizeof(double));
    }
    pl->nmtiles = 0;
    if (!e && mp && mp->ntiles > 0) {
        pl->nmtiles = mp->ntiles; pl->nsplit = mp->nsplit; pl->merge_items = mp->items; pl->merge_rows = mp->rows; pl->merge_nnz = mp->nnz;
        e = cudaMalloc((void **) &pl->d_mtiles, (size_t) mp->ntiles * sizeof(*mp->tiles));
        if (!e) e = cudaMemcpy(pl->d_mtiles, mp->tiles, (size_t) mp->ntiles * sizeof(*mp->tiles), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_msplit, (size_t) (mp->nsplit > 0 ? mp->nsplit : 1) * sizeof(*mp->split));
        if (!e && mp->nsplit > 0) e = cudaMemcpy(pl->d_msplit, mp->split, (size_t) mp->nsplit * sizeof(*mp->split), cudaMemcpyHostToDevice);
        if (!e) e = cudaMalloc((void **) &pl->d_mpart, 2 * (size_t) mp->ntiles * sizeof(double));
        if (!e) e = cudaMemset(pl->d_mpart, 0, 2 * (size_t) mp->ntiles * sizeof(double));
    }
    free(tiles); free(longrows); free(medrows);
    CU(e);
    return ACG_SUCCESS;
}

/* ext.h: row-pattern dictionary of a CSR matrix, host only (compress.c) */
int acgb200_patterns_host(int nrows, const int64_t *rowptr, const int *colidx, int max_entries,
                          int *npat, int *nentries, int *patptr, int *patoff, unsigned short *patid, int64_t *nmatched)
{
    struct acgb200_patterns pat;
    int err = acgb200_patterns_build(nrows, rowptr, colidx, max_entries, &pat);
    if (err) return err;
    *npat = pat.npat; *nentries = pat.nentries; *nmatched = pat.nrows_matched;
    memcpy(patptr, pat.patptr, ((size_t) pat.npat + 1) * sizeof(int));
    memcpy(patoff, pat.patoff, (size_t) pat.nentries * sizeof(int));
    memcpy(patid, pat.patid, (size_t) nrows * sizeof(unsigned short));
    acgb200_patterns_free(&pat);
    return ACG_SUCCESS;
}

/* ext.h: the merge-path tile plan, host only (mergeplan.c) */
int acgb200_merge_plan_host(int hi, const int64_t *rowptr, int items, int *tiles4, int maxtiles,
                            int *split3, int *ntiles, int *nsplit)
{
    struct acgb200_mergeplan mp;
    int err = acgb200_merge_plan(hi, rowptr, items, &mp);
    if (!err && mp.ntiles > maxtiles) err = ACG_ERR_NO_BUFFER_SPACE;
    if (!err) {
        for (int t = 0; t < mp.ntiles; t++) {
            tiles4[4 * t] = mp.tiles[t].r0; tiles4[4 * t + 1] = mp.tiles[t].nre;
            tiles4[4 * t + 2] = mp.tiles[t].k0; tiles4[4 * t + 3] = mp.tiles[t].nnz;
        }
        for (int i = 0; i < mp.nsplit; i++) {
            split3[3 * i] = mp.split[i].row; split3[3 * i + 1] = mp.split[i].ta; split3[3 * i + 2] = mp.split[i].tb;
        }
        *ntiles = mp.ntiles; *nsplit = mp.nsplit;
    }
    acgb200_mergeplan_free(&mp);
    return err;
}

/* ext.h: the pattern-slice plan of a CSR matrix, host only (slices.c) */
int acgb200_slices_host(int nrows, int cover_hi, const int64_t *rowptr, const int *colidx,
                        int *slices4, int maxslices, unsigned char *covered, int64_t *totals6, int *spatoff,
                        unsigned short *patid, int64_t *exc2)
{
    struct acgb200_patterns pat;
    struct acgb200_sliceplan sp;
    int err = acgb200_patterns_build(nrows, rowptr, colidx, 4096, &pat);
    if (err) return err;
    err = acgb200_slices_plan(nrows, cover_hi, rowptr, &pat, &sp);
    if (!err && sp.nslices > maxslices) err = ACG_ERR_NO_BUFFER_SPACE;
    if (!err) {
        for (int i = 0; i < sp.nslices; i++) {
            slices4[4 * i] = sp.slices[i].row0; slices4[4 * i + 1] = sp.slices[i].nrows;
            slices4[4 * i + 2] = sp.slices[i].len; slices4[4 * i + 3] = sp.slices[i].vblk;
        }
        memcpy(covered, sp.covered, (size_t) ((nrows + 31) / 32));
        totals6[0] = sp.nslices; totals6[1] = sp.blocks; totals6[2] = sp.nnz; totals6[3] = sp.rows;
        totals6[4] = sp.lpad; totals6[5] = sp.npat;
        if (sp.nslices > 0) memcpy(spatoff, sp.spatoff, (size_t) sp.npat * (size_t) sp.lpad * sizeof(int));
        if (patid && sp.nslices > 0) memcpy(patid, sp.patid, (size_t) nrows * sizeof(*patid));
        if (exc2) { exc2[0] = sp.nexc; exc2[1] = sp.excnnz; }
    }
    acgb200_sliceplan_free(&sp);
    acgb200_patterns_free(&pat);
    return err;
}

/* ext.h: the tile plan the solver would build for a CSR row-pointer array,
 * computed on the host without touching a device.  With colidx != NULL (0-based)
 * the pattern slices are planned too, exactly as acgsolvercuda_init does under option
 * "spmv_slices": the tiles then skip the rows of covered slices (acgb200_slices_host
 * returns the slices themselves). */
int acgb200_spmv_plan_host(int nrows, const int64_t *rowptr, struct acgb200_info *info,
                           int *tiles4, int maxtiles, int *longrows, int maxlong)
{
    return acgb200_spmv_plan_host2(nrows, rowptr, NULL, info, tiles4, maxtiles, longrows, maxlong);
}

int acgb200_spmv_plan_host2(int nrows, const int64_t *rowptr, const int *colidx, struct acgb200_info *info,
                            int *tiles4, int maxtiles, int *longrows, int maxlong)
{
    struct acgb200_spmvplan pl;
    memset(&pl, 0, sizeof(pl));
    int64_t maxlen = 0;
    for (int i = 0; i < nrows; i++) if (rowptr[i + 1] - rowptr[i] > maxlen) maxlen = rowptr[i + 1] - rowptr[i];
    acgb200_spmv_choose(&pl, nrows, rowptr[nrows] - rowptr[0], maxlen);
    cfg_load();
    if (cfg.spmv_lanes > 0) pl.lanes_per_row = cfg.spmv_lanes;
    if (cfg.spmv_nnz_cap > 0) pl.nnz_cap = cfg.spmv_nnz_cap;
    if (cfg.spmv_rows_cap > 0) pl.rows_cap = cfg.spmv_rows_cap;
    pl.med_thr = cfg.spmv_medium;
    struct acgb200_tile *tiles = malloc(((size_t) nrows + 1) * sizeof(*tiles));
    int *lr = malloc(((size_t) nrows + 1) * sizeof(*lr));
    int *mr = malloc(((size_t) nrows + 1) * sizeof(*mr));
    if (!tiles || !lr || !mr) { free(tiles); free(lr); free(mr); return ACG_ERR_ERRNO; }
    int nt = 0, nl = 0, nm = 0, nsl = 0, slrows = 0;
    int err = ACG_SUCCESS;
    struct acgb200_patterns pat;
    struct acgb200_sliceplan sp;
    memset(&pat, 0, sizeof(pat)); memset(&sp, 0, sizeof(sp));
    if (colidx && cfg.spmv_slices) {
        err = acgb200_patterns_build(nrows, rowptr, colidx, 4096, &pat);
        if (!err && pat.npat > 0) err = acgb200_slices_plan(nrows, nrows, rowptr, &pat, &sp);
        nsl = sp.nslices; slrows = sp.rows;
    }
    if (!err) err = cut_tiles(&pl, rowptr, sp.nslices > 0 ? sp.covered : NULL, 0, tiles, &nt, lr, &nl, mr, &nm);
    acgb200_sliceplan_free(&sp);
    acgb200_patterns_free(&pat);
    if (!err && (nt > maxtiles || nl > maxlong)) err = ACG_ERR_NO_BUFFER_SPACE;
    if (!err) {
        for (int t = 0; t < nt; t++) {
            tiles4[4 * t] = tiles[t].row_begin; tiles4[4 * t + 1] = tiles[t].nrows;
            tiles4[4 * t + 2] = tiles[t].k_al; tiles4[4 * t + 3] = tiles[t].nnz_al;
        }
        memcpy(longrows, lr, (size_t) nl * sizeof(int));
        memset(info, 0, sizeof(*info));
        info->spmv_lanes_per_row = pl.lanes_per_row; info->spmv_rows_cap = pl.rows_cap;
        info->spmv_nnz_cap = pl.nnz_cap; info->spmv_stages = pl.nstages;
        info->spmv_ntiles = nt; info->spmv_nlong = nl; info->spmv_slices = nsl; info->spmv_slice_rows = slrows;
        info->spmv_nmedium = nm;
    }
    free(tiles); free(lr); free(mr);
    return err;
}

/* upload a host int64 row-pointer array narrowed to int32 (acg/cgcuda.c:262-272)
 * with `pad` trailing entries repeating the last value */
static int upload_rowptr(int **d_out, const int64_t *rp, int64_t n, int pad, int *errcode)
{
    int *tmp = malloc(((size_t) n + 1 + (size_t) pad) * sizeof(*tmp));
    if (!tmp) return ACG_ERR_ERRNO;
    for (int64_t i = 0; i <= n; i++) {
        if (rp[i] > INT32_MAX) { free(tmp); return ACG_ERR_INDEX_OUT_OF_BOUNDS; }
        tmp[i] = (int) rp[i];
    }
    for (int i = 1; i <= pad; i++) tmp[n + i] = tmp[n];
    cudaError_t e = cudaMalloc((void **) d_out, ((size_t) n + 1 + (size_t) pad) * sizeof(*tmp));
    if (!e) e = cudaMemcpy(*d_out, tmp, ((size_t) n + 1 + (size_t) pad) * sizeof(*tmp), cudaMemcpyHostToDevice);
    free(tmp);
    if (e) { if (errcode) *errcode = (int) e; return ACG_ERR_CUDA; }
    return ACG_SUCCESS;
}

static int upload_block(int **d_col, double **d_val, const acgidx_t *col, const double *val,
                        int64_t nnz, int base, int pad, int *errcode)
{
    const size_t cap = (size_t) nnz + (size_t) pad;
    CU(cudaMalloc((void **) d_col, cap * sizeof(int)));
    CU(cudaMalloc((void **) d_val, cap * sizeof(double)));
    CU(cudaMemset(*d_col + nnz, 0, (size_t) pad * sizeof(int)));
    CU(cudaMemset(*d_val + nnz, 0, (size_t) pad * sizeof(double)));
    if (nnz > 0) {
        if (base == 0) {
            CU(cudaMemcpy(*d_col, col, (size_t) nnz * sizeof(int), cudaMemcpyHostToDevice));
        } else {
            int *tmp = malloc((size_t) nnz * sizeof(int));
            if (!tmp) return ACG_ERR_ERRNO;
            for (int64_t k = 0; k < nnz; k++) tmp[k] = col[k] - base;
            cudaError_t e = cudaMemcpy(*d_col, tmp, (size_t) nnz * sizeof(int), cudaMemcpyHostToDevice);
            free(tmp);
            CU(e);
        }
        CU(cudaMemcpy(*d_val, val, (size_t) nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    return ACG_SUCCESS;
}

static int init_impl(struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A, const struct acgcomm *comm);

int acgsolvercuda_init(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    cublasHandle_t cublas, cusparseHandle_t cusparse, const struct acgcomm *comm)
{
    (void) cublas; (void) cusparse;
    memset(cg, 0, sizeof(*cg));
    cfg_load();
    /* without full storage (no acgsymcsrmatrix_dsymv_init call) the packed triangle is expanded on the device */
    if (!A->frowptr && !A->rowptr) return ACG_ERR_INVALID_VALUE;
    if (A->frowptr && (!A->fcolidx || !A->fa)) return ACG_ERR_INVALID_VALUE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ACG_ERR_CUDA;   /* no CPU fallback */
    int err = init_impl(cg, A, comm);
    if (err) acgsolvercuda_free(cg);       /* releases whatever was set up before the failure */
    return err;
}

static int init_impl(struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A, const struct acgcomm *comm)
{
    int errcode_ = 0, *errcode = &errcode_;
    int commsize = 1;
    struct priv *pv = calloc(1, sizeof(*pv));
    if (!pv) return ACG_ERR_ERRNO;
    pv->key = cg;
    priv_add(pv);

    /* host-side work vectors, as the reference keeps them (acg/cgcuda.c:145-153) */
    OK(acgsymcsrmatrix_vector(A, &cg->r)); acgvector_setzero(&cg->r);
    OK(acgsymcsrmatrix_vector(A, &cg->p)); acgvector_setzero(&cg->p);
    OK(acgsymcsrmatrix_vector(A, &cg->t)); acgvector_setzero(&cg->t);

    cg->halo = calloc(1, sizeof(*cg->halo));            /* zeroed: acgsolvercuda_free may run before they are filled */
    cg->haloexchange = calloc(1, sizeof(*cg->haloexchange));
    if (!cg->halo || !cg->haloexchange) return ACG_ERR_ERRNO;
    OK(acgsymcsrmatrix_halo(A, cg->halo));
    {
        /* communication kernels are tiny and latency-critical: give their streams
         * the highest priority so they are scheduled ahead of SpMV CTAs */
        int plo = 0, phi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&plo, &phi));
        CU(cudaStreamCreateWithPriority(&pv->stream, cudaStreamNonBlocking, plo));
        CU(cudaStreamCreateWithPriority(&pv->commstream, cudaStreamNonBlocking, phi));
        CU(cudaStreamCreateWithPriority(&pv->redstream, cudaStreamNonBlocking, phi));
    }
    CU(cudaEventCreateWithFlags(&pv->ev_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&pv->ev_halo, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&pv->ev_red, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&pv->ev_poll[i], cudaEventDisableTiming));
    CU(cudaEventCreate(&pv->ev_t0));
    CU(cudaEventCreate(&pv->ev_t1));
    OK(acghaloexchange_init_cuda(cg->haloexchange, cg->halo, ACG_DOUBLE, ACG_DOUBLE, comm, pv->commstream));
    cg->use_nvshmem = 0;
    if (comm && comm->type == acgcomm_nvshmem) return ACG_ERR_NVSHMEM_NOT_SUPPORTED;
    {
        /* a second communicator lets the pipelined allreduce run concurrently
         * with the halo exchange (operations on one NCCL communicator are
         * serialised).  Collective: every rank calls acgsolvercuda_init. */
        OK(acgcomm_size(comm, &commsize));
        if (cfg.redstream && commsize > 1 && comm->type == acgcomm_nccl) {
            int rank = 0;
            OK(acgcomm_rank(comm, &rank));
            ncclComm_t dup = NULL;
            ncclResult_t r = ncclCommSplit(comm->ncclcomm, 0, rank, &dup, NULL);
            if (r != ncclSuccess) { *errcode = (int) r; return ACG_ERR_NCCL; }
            OK(acgcomm_init_nccl(&pv->redcomm, dup, errcode));
            pv->have_redcomm = 1;
        }
        /* peer-memory exchange; all ranks must agree on whether it is usable */
        if (cfg.p2p && commsize > 1 && comm->type == acgcomm_nccl && pv->have_redcomm) {
            int perr = 0;
            int ok = acgb200_p2p_init(&pv->p2p, cg->halo, A->borderrowoffset, A->nborderrows, comm, pv->stream, &perr) == ACG_SUCCESS;
            if (!ok) (void) cudaGetLastError();
            int *d_ok = NULL, allok = 0;
            CU(cudaMalloc((void **) &d_ok, sizeof(int)));
            { cudaError_t ce0 = cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice); if (ce0) { cudaFree(d_ok); CU(ce0); } }
            ncclResult_t r = ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm->ncclcomm, pv->stream);
            if (r != ncclSuccess) { cudaFree(d_ok); *errcode = (int) r; return ACG_ERR_NCCL; }
            cudaError_t ce = cudaMemcpyAsync(&allok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, pv->stream);
            if (!ce) ce = cudaStreamSynchronize(pv->stream);
            cudaFree(d_ok);
            CU(ce);
            if (ok) {
                pv->p2p.h_desc.fuse = cfg.p2p_fuse;
                CU(cudaMemcpy(&pv->p2p.d_desc->fuse, &pv->p2p.h_desc.fuse, sizeof(int), cudaMemcpyHostToDevice));
            }
            if (!allok) {
                if (getenv("ACGB200_VERBOSE")) fprintf(stderr, "acgb200: peer-memory exchange unavailable (cuda error %d), using NCCL\n", perr);
                pv->p2p.enabled = 0;
            }
        }
    }

    pv->nowned = A->nownedrows; pv->ninner = A->ninnerrows; pv->nborder = A->nborderrows;
    pv->nghost = A->nghostrows; pv->borderoff = A->borderrowoffset;
    pv->nvec = cg->r.num_nonzeros;

    /* device scalars and control block */
    CU(cudaMalloc((void **) &pv->d_st, sizeof(*pv->d_st)));
    CU(cudaMemset(pv->d_st, 0, sizeof(*pv->d_st)));
    CU(cudaMallocHost((void **) &pv->h_ctrl, 2 * sizeof(*pv->h_ctrl)));
    CU(cudaMallocHost((void **) &pv->h_st, sizeof(*pv->h_st)));

    /* device vectors: owned + ghost entries, padded to an even count */
    const size_t vbytes = ((size_t) pv->nvec + 2) * sizeof(double);
    CU(cudaMalloc((void **) &cg->d_r, vbytes)); CU(cudaMemset(cg->d_r, 0, vbytes));
    pv->d_r_id = cg->d_r;
    CU(cudaMalloc((void **) &cg->d_p, vbytes)); CU(cudaMemset(cg->d_p, 0, vbytes));
    CU(cudaMalloc((void **) &cg->d_t, vbytes)); CU(cudaMemset(cg->d_t, 0, vbytes));

    /* local block (rows [0,nowned) of the full storage) and border x ghost block */
    const int64_t *frp = A->frowptr;          /* host row pointers / columns of the local block, for the plans */
    const acgidx_t *fcol = A->fcolidx;
    int64_t *frp_tmp = NULL;
    acgidx_t *fcol_tmp = NULL;
    if (A->frowptr) {
        OK(upload_rowptr(&cg->d_rowptr, A->frowptr, A->nprows, 8, errcode));
        OK(upload_block(&cg->d_colidx, &cg->d_a, A->fcolidx, A->fa, A->fnpnzs, A->rowidxbase, 16, errcode));
        OK(upload_rowptr(&cg->d_orowptr, A->orowptr, (int64_t) A->nborderrows + A->nghostrows, 8, errcode));
        OK(upload_block(&cg->d_ocolidx, &cg->d_oa, A->ocolidx, A->oa, A->onpnzs, A->rowidxbase, 16, errcode));
        pv->fnnz = A->fnpnzs; pv->onnz = A->onpnzs;
    } else {
        /* the caller did not build full storage: upload the packed triangle (half the bytes) and mirror it
         * on the device (expand.cu) -- the arrays are those acgsymcsrmatrix_dsymv_init(A, 0) would have made */
        struct acgb200_expanded ex;
        OK(acgb200_expand_upload(A, 0.0, 8, 16, &ex, pv->stream, errcode));
        cg->d_rowptr = ex.d_rowptr; cg->d_colidx = ex.d_colidx; cg->d_a = ex.d_a;
        cg->d_orowptr = ex.d_orowptr; cg->d_ocolidx = ex.d_ocolidx; cg->d_oa = ex.d_oa;
        pv->fnnz = ex.fnnz; pv->onnz = ex.onnz;
        pv->device_expanded = 1;
        /* the planners work on the host: row pointers always, columns only for the pattern dictionary */
        int *t = malloc(((size_t) A->nprows + 1) * sizeof(*t));
        frp_tmp = malloc(((size_t) A->nprows + 1) * sizeof(*frp_tmp));
        if (!t || !frp_tmp) { free(t); free(frp_tmp); return ACG_ERR_ERRNO; }
        cudaError_t ce = cudaMemcpy(t, cg->d_rowptr, ((size_t) A->nprows + 1) * sizeof(int), cudaMemcpyDeviceToHost);
        for (int64_t i = 0; i <= A->nprows; i++) frp_tmp[i] = t[i];
        free(t);
        if (!ce && cfg.spmv_slices && ex.fnnz > 0) {
            fcol_tmp = malloc((size_t) ex.fnnz * sizeof(*fcol_tmp));
            if (!fcol_tmp) { free(frp_tmp); return ACG_ERR_ERRNO; }
            ce = cudaMemcpy(fcol_tmp, cg->d_colidx, (size_t) ex.fnnz * sizeof(int), cudaMemcpyDeviceToHost);
        }
        if (ce) { free(frp_tmp); free(fcol_tmp); *errcode = (int) ce; return ACG_ERR_CUDA; }
        frp = frp_tmp; fcol = fcol_tmp;
    }
    const int colbase = A->frowptr ? A->rowidxbase : 0;     /* device-expanded columns are 0-based */

    /* SpMV tile plan */
    int64_t maxlen = 0;
    for (acgidx_t i = 0; i < A->nownedrows; i++) {
        const int64_t len = frp[i + 1] - frp[i];
        if (len > maxlen) maxlen = len;
    }
    acgb200_spmv_choose(&pv->plan, A->nownedrows, frp[A->nownedrows], maxlen);
    if (cfg.spmv_lanes > 0) pv->plan.lanes_per_row = cfg.spmv_lanes;
    if (cfg.spmv_nnz_cap > 0) pv->plan.nnz_cap = cfg.spmv_nnz_cap;
    if (cfg.spmv_rows_cap > 0) pv->plan.rows_cap = cfg.spmv_rows_cap;
    if (cfg.spmv_stages > 0) pv->plan.nstages = cfg.spmv_stages > 8 ? 8 : cfg.spmv_stages;
    if (cfg.spmv_threads > 0) pv->plan.threads = cfg.spmv_threads;
    if (cfg.spmv_unroll > 0) pv->plan.unroll = cfg.spmv_unroll;
    pv->plan.med_thr = cfg.spmv_medium;
    {
        struct acgb200_patterns pat;
        struct acgb200_sliceplan sp;
        memset(&pat, 0, sizeof(pat));
        memset(&sp, 0, sizeof(sp));
        int err = ACG_SUCCESS;
        if (cfg.spmv_slices && colbase == 0 && fcol)
            err = acgb200_patterns_build(A->nownedrows, frp, fcol, 4096, &pat);
        if (!err && cfg.spmv_slices && pat.npat > 0) {
            /* between GPUs the border rows stay with the tile kernel, which adds the border x ghost block */
            const int cover_hi = (commsize > 1 || A->nghostrows > 0) ? A->borderrowoffset : A->nownedrows;
            err = acgb200_slices_plan(A->nownedrows, cover_hi, frp, &pat, &sp);
        }
        /* irregular row lengths (power-law matrices): merge-path tiles instead of row-aligned ones for the rows
         * the slices did not take.  "Irregular": the longest row is far above the mean. */
        struct acgb200_mergeplan mp;
        memset(&mp, 0, sizeof(mp));
        if (!err && sp.nslices == 0 && cfg.spmv_merge != 0) {
            const int hi = (commsize > 1 || A->nghostrows > 0) ? A->borderrowoffset : A->nownedrows;
            const double avg = hi > 0 ? (double) frp[hi] / hi : 0.0;
            int64_t mx = 0;
            for (int i = 0; i < hi; i++) if (frp[i + 1] - frp[i] > mx) mx = frp[i + 1] - frp[i];
            if (hi >= 1024 && (cfg.spmv_merge > 0 || (double) mx > 8.0 * avg + 64.0))
                /* 2048 merged items per tile: measured on R-MAT 20 M (profiles/r02/c_ab_rmat20m.log), 4.32 ms against
                 * 6.33 ms with 1024 and 4.74 ms with 4096 */
                err = acgb200_merge_plan(hi, frp, cfg.merge_items > 0 ? cfg.merge_items : 2048, &mp);
        }
        if (!err) err = build_tiles(&pv->plan, frp, &sp, pat.patid, &mp, errcode);
        if (!err && pv->plan.nmtiles > 0) {
            pv->plan.merge_threads = cfg.merge_threads > 0 ? cfg.merge_threads : 128;
            pv->plan.merge_stages = cfg.merge_stages >= 1 && cfg.merge_stages <= SPMV_MAX_STAGES_HOST ? cfg.merge_stages : 2;
            pv->plan.merge_max_ctas = cfg.merge_max_ctas;
        }
        acgb200_mergeplan_free(&mp);
        if (!err && pv->plan.nslices > 0) {
            const int d = sp.domlen;
            pv->plan.slice_ub = cfg.slice_ub > 0 ? cfg.slice_ub : (d % 9 == 0 ? 9 : d % 7 == 0 ? 7 : d % 8 == 0 ? 8 : d % 5 == 0 ? 5 : 8);
            pv->plan.slice_threads = cfg.slice_threads > 0 ? cfg.slice_threads : 128;
            pv->plan.slice_max_ctas = cfg.slice_max_ctas;
        }
        acgb200_sliceplan_free(&sp);
        acgb200_patterns_free(&pat);
        free(frp_tmp); free(fcol_tmp);
        if (err) return err;
    }
    pv->plan.max_ctas_per_sm = cfg.spmv_max_ctas;
    KL(acgb200_spmv_configure(&pv->plan));
    if (pv->plan.nslices > 0) {
        /* slice-major copy of the covered rows' values, made on the device from the CSR values */
        KL(acgb200_slices_fill(&pv->plan, cg->d_rowptr, cg->d_a, pv->stream));
        CU(cudaStreamSynchronize(pv->stream));
    }
    return ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* building blocks of one solve                                              */
/* ------------------------------------------------------------------------ */

struct solvectx {
    struct acgsolvercuda *cg;
    struct priv *pv;
    const struct acgcomm *comm;
    int multi, tag;
    int *errcode;
    double *d_b, *d_x;
    int launches;
    int capturing;            /* inside cudaStreamBeginCapture: no profiling marks */
    int p2p;                  /* loop exchanges go through peer memory */
};

static int evpool_reserve(struct evpool *p, int n)
{
    if (n <= p->cap) return 0;
    cudaEvent_t *e = realloc(p->ev, (size_t) n * sizeof(*e));
    if (!e) return ACG_ERR_ERRNO;
    p->ev = e;
    for (int i = p->cap; i < n; i++) if (cudaEventCreate(&p->ev[i])) return ACG_ERR_CUDA;
    p->cap = n;
    return 0;
}

static void prof_mark(struct solvectx *c, struct evpool *p)
{
    if (!cfg.profile || c->capturing || p->n >= p->cap) return;
    cudaEventRecord(p->ev[p->n++], c->pv->stream);
}

static double evpool_sum_ms(struct evpool *p)
{
    double tot = 0;
    for (int i = 0; i + 1 < p->n; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p->ev[i], p->ev[i + 1]) == cudaSuccess) tot += ms;
    }
    return tot;
}

/* y = A x (or r = b - A x) over owned rows, including the halo exchange of x
 * and the border x ghost block when the matrix is distributed.  `acc` gets the
 * fused dot.  `gated`: take part in the device-side iteration control. */
static int apply_A(struct solvectx *c, const double *x_ro, double *x_halo, double *y, const double *b,
                   int mode, double *acc, int gated, int housekeeping, int warmup, int pub_ch)
{
    struct acgsolvercuda *cg = c->cg;
    struct priv *pv = c->pv;
    int *errcode = c->errcode;
    /* loop iterations exchange ghosts through peer memory (already pushed by the
     * producer of x); set-up products (x0, r0) use the NCCL exchange */
    const int peer = c->multi && c->p2p && gated;
    if (c->multi && !peer) {
        /* the vector was produced on the main stream */
        CU(cudaEventRecord(pv->ev_ready, pv->stream));
        CU(cudaStreamWaitEvent(pv->commstream, pv->ev_ready, 0));
        OK(acghalo_exchange_cuda_begin(cg->halo, cg->haloexchange, pv->nvec, x_halo, ACG_DOUBLE,
                                       pv->nvec, x_halo, ACG_DOUBLE, c->comm, c->tag, errcode, warmup, pv->commstream));
        c->launches += 2;
    }
    struct acgb200_spmvargs a;
    memset(&a, 0, sizeof(a));
    a.plan = &pv->plan;
    a.rowptr = cg->d_rowptr; a.colidx = cg->d_colidx; a.a = cg->d_a;
    a.x = x_ro; a.y = y; a.b = b; a.acc = acc;
    a.dotrows = (c->multi && !peer) ? pv->borderoff : pv->nowned;
    a.mode = mode;
    if (gated) { a.ctrl_in = &pv->d_st->ctrl[0]; a.ctrl_out = &pv->d_st->ctrl[1]; }
    a.st = pv->d_st; a.housekeeping = housekeeping;
    a.pub_ch = -1;
    const int fused = peer && pv->p2p.h_desc.fuse;
    if (peer && !fused) a.dotrows = pv->borderoff;
    if (fused) {
        /* one kernel: local block, then (for border rows) the border x ghost
         * block with ghosts read from the window, fused dot over all rows and,
         * if asked, publication of the dot to all ranks by the last CTA */
        a.p2p = pv->p2p.d_desc;
        a.od_rowoffset = pv->borderoff; a.od_nrows = pv->nborder;
        a.orowptr = cg->d_orowptr; a.ocolidx = cg->d_ocolidx; a.oa = cg->d_oa;
        if (pub_ch >= 0 && a.plan->nlong == 0 && a.plan->nmed == 0) a.pub_ch = pub_ch;
    }
    prof_mark(c, &pv->gemv);
    KL(acgb200_spmv_launch(&a, pv->stream));
    {
        /* kernels acgb200_spmv_launch issues for this plan (kernels.cu): slices | merge tiles (+ fix-up), the tile
         * kernel unless the first one forwards the control word itself, medium rows, long rows (two kernels) */
        const struct acgb200_spmvplan *pl = a.plan;
        const int first = pl->nslices > 0 || pl->nmtiles > 0;
        const int tilek = pl->ntiles > 0 || (a.ctrl_in && !(first && pl->ntiles == 0 && !a.p2p));
        c->launches += (pl->nslices > 0) + (pl->nmtiles > 0) + (pl->nmtiles > 0 && pl->nsplit > 0) + tilek
                       + (pl->nlong > 0 ? 2 : 0) + (pl->nmed > 0 ? 1 : 0);
    }
    if (c->multi && !fused) {
        if (!peer) {
            OK(acghalo_exchange_cuda_end(cg->halo, cg->haloexchange, pv->nvec, x_halo, ACG_DOUBLE,
                                         pv->nvec, x_halo, ACG_DOUBLE, c->comm, c->tag, errcode, warmup, pv->commstream));
            CU(cudaEventRecord(pv->ev_halo, pv->commstream));
            CU(cudaStreamWaitEvent(pv->stream, pv->ev_halo, 0));
        }
        struct acgb200_offdiagargs o;
        memset(&o, 0, sizeof(o));
        o.nrows = pv->nborder; o.rowoffset = pv->borderoff;
        o.orowptr = cg->d_orowptr; o.ocolidx = cg->d_ocolidx; o.oa = cg->d_oa;
        o.x = x_ro; o.y = y; o.acc = acc;
        o.minus = (mode == SPMV_R_B_AX);
        o.dotkind = !acc ? 0 : (mode == SPMV_R_B_AX ? 2 : (mode == SPMV_Y_AX_DOT ? 1 : 0));
        if (gated) o.ctrl_in = &pv->d_st->ctrl[1];
        o.st = pv->d_st;
        o.p2p = peer ? pv->p2p.d_desc : NULL; o.p2p_iter_override = -1;
        KL(acgb200_offdiag_launch(&o, pv->stream));
        c->launches += (pv->nborder > 0);
    }
    prof_mark(c, &pv->gemv);
    return ACG_SUCCESS;
}

/* push the border entries of `vec` (and/or reduction partials) into the peers'
 * windows; see struct acgb200_postargs */
static int post(struct solvectx *c, int ctrl_slot, int iter_override, const double *vec,
                int ch, const double *redbase, int stride, int count, int par_off, int seq_off)
{
    struct priv *pv = c->pv;
    int *errcode = c->errcode;
    struct acgb200_postargs a;
    memset(&a, 0, sizeof(a));
    a.p2p = pv->p2p.d_desc;
    a.cin = &pv->d_st->ctrl[ctrl_slot]; a.st = pv->d_st;
    a.iter_override = iter_override;
    a.vec = vec; a.sendbufidx = (const int *) c->cg->haloexchange->d_sendbufidx;
    a.ch = ch; a.redbase = redbase; a.redstride = stride; a.redcount = count; a.par_off = par_off; a.seq_off = seq_off;
    KL(acgb200_comm_post(&a, pv->stream));
    c->launches += 1;
    return ACG_SUCCESS;
}

static int allreduce(struct solvectx *c, const double *src, double *dst, int count)
{
    if (!c->multi) return ACG_SUCCESS;
    return acgcomm_allreduce(src, dst, count, ACG_DOUBLE, ACG_SUM, c->pv->stream, c->comm, c->errcode);
}

/* global sum of a device-side reduction result, brought to the host */
static int reduce_to_host(struct solvectx *c, double *loc, double *glob, int count, double *host)
{
    int *errcode = c->errcode;
    OK(allreduce(c, loc, glob, count));
    CU(cudaMemcpyAsync(host, c->multi ? glob : loc, (size_t) count * sizeof(double), cudaMemcpyDeviceToHost, c->pv->stream));
    CU(cudaStreamSynchronize(c->pv->stream));
    return ACG_SUCCESS;
}

static int solve_begin(struct solvectx *c, struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
                       const struct acgvector *b, struct acgvector *x, double diffatol, double diffrtol,
                       struct acgcomm *comm, int tag, int *errcode)
{
    memset(c, 0, sizeof(*c));
    if (b->size < A->nrows || x->size < A->nrows || cg->r.size < A->nrows ||
        cg->p.size < A->nrows || cg->t.size < A->nrows) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    if (diffatol > 0 || diffrtol > 0) return ACG_ERR_NOT_SUPPORTED;     /* acg/cgcuda.c:424 */
    struct priv *pv = priv_of(cg);
    if (!pv) return ACG_ERR_INVALID_VALUE;
    if (b->num_nonzeros < pv->nowned || x->num_nonzeros < pv->nvec) return ACG_ERR_VECTOR_INCOMPATIBLE_SIZE;
    int commsize = 1;
    OK(acgcomm_size(comm, &commsize));
    c->cg = cg; c->pv = pv; c->comm = comm; c->multi = commsize > 1; c->tag = tag; c->errcode = errcode;
    if (c->multi && comm->type != acgcomm_nccl)
        return comm->type == acgcomm_mpi ? ACG_ERR_MPI_NOT_SUPPORTED : ACG_ERR_NVSHMEM_NOT_SUPPORTED;
    /* b and x0 to the device (acg/cgcuda.c:484-493).  The reference allocates and
     * frees the two device vectors in every solve; here they are kept with the
     * solver (their addresses are baked into the replayed CUDA graphs). */
    const double tb = wall();
    const size_t vbytes = ((size_t) pv->nvec + 2) * sizeof(double);
    if (!pv->d_b) {
        CU(cudaMalloc((void **) &pv->d_b, vbytes));
        CU(cudaMalloc((void **) &pv->d_x, vbytes));
        CU(cudaMemsetAsync(pv->d_b, 0, vbytes, pv->stream));
        CU(cudaMemsetAsync(pv->d_x, 0, vbytes, pv->stream));
    }
    c->d_b = pv->d_b; c->d_x = pv->d_x;
    /* a caller may hand over longer vectors than this part's owned + ghost entries: only those are used */
    const size_t nb = (size_t) (b->num_nonzeros < pv->nvec ? b->num_nonzeros : pv->nvec);
    const size_t nx = (size_t) (x->num_nonzeros < pv->nvec ? x->num_nonzeros : pv->nvec);
    CU(cudaMemcpyAsync(c->d_b, b->x, nb * sizeof(double), cudaMemcpyHostToDevice, pv->stream));
    CU(cudaMemcpyAsync(c->d_x, x->x, nx * sizeof(double), cudaMemcpyHostToDevice, pv->stream));
    CU(cudaStreamSynchronize(pv->stream));
    pv->last_h2d_ms = 1e3 * (wall() - tb);
    return ACG_SUCCESS;
}

static int solve_end(struct solvectx *c, struct acgvector *x, int status)
{
    int *errcode = c->errcode;
    const double te = wall();
    const size_t nx = (size_t) (x->num_nonzeros < c->pv->nvec ? x->num_nonzeros : c->pv->nvec);
    CU(cudaMemcpyAsync(x->x, c->d_x, nx * sizeof(double), cudaMemcpyDeviceToHost, c->pv->stream));
    CU(cudaStreamSynchronize(c->pv->stream));
    c->pv->last_d2h_ms = 1e3 * (wall() - te);
    CU(cudaGetLastError());
    return status;
}

static int push_state(struct solvectx *c, const struct acgb200_devstate *st)
{
    int *errcode = c->errcode;
    /* everything queued before must be done with the staging buffer and with
     * the device state it is about to replace */
    CU(cudaStreamSynchronize(c->pv->stream));
    *c->pv->h_st = *st;
    CU(cudaMemcpyAsync(c->pv->d_st, c->pv->h_st, sizeof(*st), cudaMemcpyHostToDevice, c->pv->stream));
    CU(cudaStreamSynchronize(c->pv->stream));
    return ACG_SUCCESS;
}

static int pull_state(struct solvectx *c, struct acgb200_devstate *st)
{
    int *errcode = c->errcode;
    CU(cudaMemcpyAsync(c->pv->h_st, c->pv->d_st, sizeof(*st), cudaMemcpyDeviceToHost, c->pv->stream));
    CU(cudaStreamSynchronize(c->pv->stream));
    *st = *c->pv->h_st;
    return ACG_SUCCESS;
}

static double threshold(double atol, double rtol, double r0nrm2)
{
    /* acg/cgcuda.c:833,:1011-1012: converged if ||r|| < atol or ||r|| < rtol*||r0||,
     * each only when positive -- i.e. ||r|| < max of the enabled thresholds */
    double t = 0;
    if (atol > 0) t = atol;
    if (rtol * r0nrm2 > t) t = rtol * r0nrm2;
    return t;
}

/* Capture iterations (parity 0, parity 1) of `issue` into a graph.  All
 * pointers the kernels and NCCL calls use are fixed for the life of the solver
 * and the iteration index only enters through its parity, so one two-iteration
 * graph serves the whole solve (and later solves). */
/* everything besides the (fixed) pointers that decides which kernels and NCCL calls an iteration issues:
 * a cached graph is only replayed for the configuration it was captured with (options can change between
 * two solves on one solver) */
static int graph_signature(const struct solvectx *c)
{
    return 1 | (c->multi ? 2 : 0) | (c->p2p ? 4 : 0) | (c->pv->p2p.h_desc.fuse ? 8 : 0) | (cfg.pdl ? 16 : 0)
           | (cfg.redstream ? 32 : 0) | (c->pv->have_redcomm ? 64 : 0);
}

static int capture_pair(struct solvectx *c, int kind, int (*issue)(struct solvectx *, int))
{
    struct priv *pv = c->pv;
    int *errcode = c->errcode;
    if (pv->graph[kind] && pv->graph_sig[kind] == graph_signature(c)) return ACG_SUCCESS;
    if (pv->graph[kind]) { cudaGraphExecDestroy(pv->graph[kind]); pv->graph[kind] = NULL; }
    const int before = c->launches;
    cudaGraph_t g = NULL;
    c->capturing = 1;
    CU(cudaStreamBeginCapture(pv->stream, cudaStreamCaptureModeThreadLocal));
    int err = issue(c, 2);
    if (!err) err = issue(c, 3);
    cudaError_t e = cudaStreamEndCapture(pv->stream, &g);
    c->capturing = 0;
    pv->graph_launches[kind] = c->launches - before;
    c->launches = before;
    if (err) { if (g) cudaGraphDestroy(g); return err; }
    CU(e);
    e = cudaGraphInstantiate(&pv->graph[kind], g, 0);
    cudaGraphDestroy(g);
    CU(e);
    pv->graph_sig[kind] = graph_signature(c);
    return ACG_SUCCESS;
}

/* Run `issue(c,k)` for k < maxits, polling the device control word.  The
 * first two iterations are issued directly (they also serve as the un-captured
 * first use of every NCCL path), the rest as replays of the two-iteration graph. */
static int iterate(struct solvectx *c, int maxits, int poll, int kind, int (*issue)(struct solvectx *, int))
{
    struct priv *pv = c->pv;
    int *errcode = c->errcode;
    int issued = 0, slot = 0, have_prev = 0, since_poll = 0;
    /* Replays are used on one GPU and with the peer-memory exchange (kernel nodes
     * only).  With the NCCL exchange the iteration also contains collectives on
     * two communicators and three streams; that combination ran in the
     * benchmarks but its test matrix is not closed yet, so it is only captured
     * on request (option "graph" = 2). */
    const int use_graph = cfg.graph && !cfg.profile && maxits >= 6 && (!c->multi || c->p2p || cfg.graph >= 2);
    while (issued < maxits) {
        if (use_graph && issued >= 2 && maxits - issued >= 2) {
            if (!pv->graph[kind] || pv->graph_sig[kind] != graph_signature(c)) OK(capture_pair(c, kind, issue));
            CU(cudaGraphLaunch(pv->graph[kind], pv->stream));
            c->launches += pv->graph_launches[kind];
            issued += 2; since_poll += 2;
        } else {
            OK(issue(c, issued));
            issued += 1; since_poll += 1;
        }
        if (!poll || (since_poll < cfg.check_every && issued < maxits)) continue;
        since_poll = 0;
        CU(cudaMemcpyAsync(&pv->h_ctrl[slot], &pv->d_st->ctrl[0], sizeof(struct acgb200_ctrl),
                           cudaMemcpyDeviceToHost, pv->stream));
        CU(cudaEventRecord(pv->ev_poll[slot], pv->stream));
        if (have_prev) {
            CU(cudaEventSynchronize(pv->ev_poll[slot ^ 1]));
            if (pv->h_ctrl[slot ^ 1].done) break;
        }
        have_prev = 1; slot ^= 1;
    }
    CU(cudaStreamSynchronize(pv->stream));
    if (c->multi) { CU(cudaStreamSynchronize(pv->commstream)); CU(cudaStreamSynchronize(pv->redstream)); }
    if (c->multi && c->p2p) {
        /* a kernel that waited in vain for a peer gave up instead of hanging the GPU */
        int gaveup = 0;
        OK(acgb200_p2p_timed_out(&pv->p2p, pv->stream, &gaveup));
        if (gaveup) { *errcode = (int) cudaErrorLaunchTimeout; return ACG_ERR_CUDA; }
    }
    return ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* classic CG                                                                */
/* ------------------------------------------------------------------------ */

static int classic_iteration(struct solvectx *c, int k)
{
    struct acgsolvercuda *cg = c->cg;
    struct priv *pv = c->pv;
    struct acgb200_devstate *st = pv->d_st;
    int *errcode = c->errcode;
    const int s = k & 1, n = pv->nowned;
    const int peer = c->multi && c->p2p;
    double *pvec = cg->d_p;
    struct acgb200_p2pdev *desc = !peer ? NULL : pv->p2p.d_desc;
    const struct acgb200_spmvplan *pl = &pv->plan;
    OK(apply_A(c, pvec, pvec, cg->d_t, NULL, SPMV_Y_AX_DOT, &st->pap_loc[s], 1, 1, 0, 0));
    if (peer) {
        /* (p,Ap) is published by the SpMV's last CTA; with long rows the dot is
         * only complete after the finishing kernel, so a separate post does it */
        if (pl->nlong > 0 || pl->nmed > 0 || !pv->p2p.h_desc.fuse) OK(post(c, 1, -1, NULL, 0, &st->pap_loc[0], 1, 1, 0, 1));
    } else OK(allreduce(c, &st->pap_loc[s], &st->pap[s], 1));
    prof_mark(c, &pv->blas);
    KL(acgb200_cg_update_r(n, st, 1, 1, c->multi, desc, cg->d_t, cg->d_r, pv->stream));
    if (!peer) OK(allreduce(c, &st->rr_loc[s ^ 1], &st->rr[s ^ 1], 1));
    else if (!pv->p2p.h_desc.fuse) OK(post(c, 1, -1, NULL, 1, &st->rr_loc[0], 1, 1, 1, 1));
    KL(acgb200_cg_update_xp(n, st, 1, 0, c->multi, desc, cg->d_r, pvec, c->d_x, pv->stream));
    prof_mark(c, &pv->blas);
    if (peer && !pv->p2p.h_desc.fuse) OK(post(c, 0, -1, pvec, -1, NULL, 0, 0, 0, 0));     /* p for the next SpMV */
    c->launches += 2 + (c->multi && !peer ? 2 : 0);
    return ACG_SUCCESS;
}

static void account_classic(struct acgsolvercuda *cg, const struct priv *pv, int nits, int multi)
{
    /* the reference's analytic counters (acg/cgcuda.c:885-1001), per iteration */
    const int64_t n = pv->nowned, nnz = pv->fnnz + pv->onnz;
    const int64_t bgemv = nnz * 12 + n * 16 + (int64_t) (pv->nborder + pv->nghost) * 8 + (int64_t) pv->nvec * 8;
    cg->ngemv += nits; cg->nflops += (int64_t) nits * 3 * nnz; cg->Bgemv += (int64_t) nits * bgemv;
    cg->ndot += nits; cg->nflops += (int64_t) nits * 2 * n; cg->Bdot += (int64_t) nits * 16 * n;
    cg->nnrm2 += nits; cg->nflops += (int64_t) nits * 2 * n; cg->Bnrm2 += (int64_t) nits * 8 * n;
    cg->naxpy += 3 * (int64_t) nits; cg->nflops += (int64_t) nits * 6 * n; cg->Baxpy += (int64_t) nits * 48 * n;
    if (multi) {
        cg->nallreduce += 2 * (int64_t) nits; cg->Ballreduce += (int64_t) nits * 16;
        cg->nhalo += nits; cg->Bhalo += (int64_t) nits * cg->halo->sendsize * 8;
        cg->nhalomsgs += (int64_t) nits * cg->halo->nrecipients;
    }
}

int acgsolvercuda_solvempi(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int tag, int *errcode,
    cublasHandle_t cublas, cusparseHandle_t cusparse, cusparseSpMVAlg_t alg)
{
    (void) cublas; (void) cusparse; (void) alg;
    int errcode_ = 0;
    if (!errcode) errcode = &errcode_;
    cfg_load();
    struct solvectx c;
    OK(solve_begin(&c, cg, A, b, x, diffatol, diffrtol, comm, tag, errcode));
    struct priv *pv = c.pv;
    struct acgb200_devstate *st = pv->d_st;
    const int n = pv->nowned;
    struct acgb200_devstate h;

    c.p2p = c.multi && pv->p2p.enabled && cfg.p2p;
    /* a rank that is ahead must not store into a peer's window while that peer's kernels of the
     * previous solve are still reading it: order the first post of this solve behind every rank */
    if (c.p2p) OK(acgcomm_barrier(pv->stream, comm, errcode));
    double *const pvec = cg->d_p;
    /* warm-up: `warmup` full iterations (every kernel, every communication path)
     * on state that is overwritten below (acg/cgcuda.c:607-705); d_r stands in
     * for x so the initial guess is untouched */
    if (warmup > 0) {
        memset(&h, 0, sizeof(h));
        h.maxits = warmup; h.rr_loc[0] = h.rr[0] = 1.0;
        OK(push_state(&c, &h));
        double *xsave = c.d_x; c.d_x = cg->d_r;
        if (c.p2p) {
            OK(acgb200_p2p_begin(&pv->p2p, warmup, pv->stream));
            OK(post(&c, 0, 0, pvec, -1, NULL, 0, 0, 0, 0));
        }
        for (int i = 0; i < warmup; i++) OK(classic_iteration(&c, i));
        c.d_x = xsave;
        KL(acgb200_dot(n, c.d_b, c.d_b, &st->tmp_loc[0], pv->stream));
        CU(cudaStreamSynchronize(pv->stream));
    }
    if (cfg.profile) {
        OK(evpool_reserve(&pv->gemv, 2 * (maxits + 2)));
        OK(evpool_reserve(&pv->blas, 2 * (maxits + 2)));
    }
    pv->gemv.n = pv->blas.n = 0;

    cg->nsolves++; cg->niterations = 0;
    cg->bnrm2 = cg->r0nrm2 = cg->rnrm2 = cg->x0nrm2 = cg->dxnrm2 = INFINITY;
    cg->maxits = maxits; cg->diffatol = diffatol; cg->diffrtol = diffrtol;
    cg->residualatol = residualatol; cg->residualrtol = residualrtol;
    OK(acgcomm_barrier(pv->stream, comm, errcode));
    CU(cudaStreamSynchronize(pv->stream));
    const double t0 = wall();
    CU(cudaEventRecord(pv->ev_t0, pv->stream));
    c.launches = 0;

    /* ||b|| (acg/cgcuda.c:727-739) */
    memset(&h, 0, sizeof(h));
    OK(push_state(&c, &h));
    double bb = 0, rr0 = 0;
    KL(acgb200_dot(n, c.d_b, c.d_b, &st->tmp_loc[0], pv->stream));
    OK(reduce_to_host(&c, &st->tmp_loc[0], &st->tmp[0], 1, &bb));
    cg->bnrm2 = sqrt(bb);
    cg->nnrm2++; cg->nflops += 2 * (int64_t) n; cg->Bnrm2 += 8 * (int64_t) n;

    /* r0 = b - A x0 with (r0,r0) folded in (acg/cgcuda.c:761-799,:819-832); p = r0 */
    OK(apply_A(&c, c.d_x, c.d_x, cg->d_r, c.d_b, SPMV_R_B_AX, &st->rr_loc[0], 0, 0, 0, -1));
    CU(cudaMemcpyAsync(pvec, cg->d_r, (size_t) n * sizeof(double), cudaMemcpyDeviceToDevice, pv->stream));
    OK(reduce_to_host(&c, &st->rr_loc[0], &st->rr[0], 1, &rr0));
    cg->rnrm2 = cg->r0nrm2 = sqrt(rr0);
    cg->ngemv++; cg->ncopy += 2; cg->nnrm2++;
    const double tol = threshold(residualatol, residualrtol, cg->r0nrm2);
    const double rtol_scaled = residualrtol * cg->r0nrm2;
    int converged = tol > 0 && cg->rnrm2 < tol;                    /* acg/cgcuda.c:836-842 */

    if (!converged && maxits > 0) {
        memset(&h, 0, sizeof(h));
        h.maxits = maxits; h.tol = tol;
        h.rr_loc[0] = h.rr[0] = rr0;
        OK(push_state(&c, &h));
        if (c.p2p) {
            /* p_0 = r_0 goes to the neighbours' windows as exchange number 0 */
            OK(acgb200_p2p_begin(&pv->p2p, maxits, pv->stream));
            OK(post(&c, 0, 0, pvec, -1, NULL, 0, 0, 0, 0));
        }
        OK(iterate(&c, maxits, tol > 0, 0, classic_iteration));
        OK(pull_state(&c, &h));
        cg->niterations = h.ctrl[0].iter;
        converged = h.ctrl[0].done;
        const double rr = converged ? h.final_rr
            : (c.multi ? h.rr[cg->niterations & 1] : h.rr_loc[cg->niterations & 1]);
        cg->rnrm2 = sqrt(rr);
        cg->ntotaliterations += cg->niterations;
    }
    CU(cudaEventRecord(pv->ev_t1, pv->stream));
    CU(cudaEventSynchronize(pv->ev_t1));
    const double t1 = wall();
    cg->tsolve += t1 - t0;
    { float ms = 0; if (cudaEventElapsedTime(&ms, pv->ev_t0, pv->ev_t1) == cudaSuccess) pv->last_solve_ms = ms; }
    account_classic(cg, pv, cg->niterations, c.multi);
    pv->last_launches = c.launches;
    if (cfg.profile) {
        pv->last_spmv_ms = evpool_sum_ms(&pv->gemv); pv->last_spmv_n = pv->gemv.n / 2;
        pv->last_blas_ms = evpool_sum_ms(&pv->blas);
        cg->tgemv += 1e-3 * pv->last_spmv_ms;
        cg->taxpy += 1e-3 * pv->last_blas_ms;
    }
    int status = ACG_SUCCESS;
    if (!converged && !(residualatol == 0 && rtol_scaled == 0)) status = ACG_ERR_NOT_CONVERGED;   /* :1099-1107 */
    return solve_end(&c, x, status);
}

int acgsolvercuda_solve(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol, int warmup)
{
    struct acgcomm null;
    memset(&null, 0, sizeof(null));
    null.type = acgcomm_null;
    int errcode = 0;
    return acgsolvercuda_solvempi(cg, A, b, x, maxits, diffatol, diffrtol, residualatol, residualrtol,
                                  warmup, &null, 0, &errcode, NULL, NULL, 0);
}

/* ------------------------------------------------------------------------ */
/* pipelined CG                                                              */
/* ------------------------------------------------------------------------ */

static int ensure_vec(struct acgvector **hv, double **dv, const struct acgsymcsrmatrix *A, int nvec, int *errcode)
{
    if (!*hv) {
        *hv = malloc(sizeof(**hv));
        if (!*hv) return ACG_ERR_ERRNO;
        OK(acgsymcsrmatrix_vector(A, *hv));
        acgvector_setzero(*hv);
    }
    if (!*dv) CU(cudaMalloc((void **) dv, ((size_t) nvec + 2) * sizeof(double)));
    return ACG_SUCCESS;
}

static int pipelined_iteration(struct solvectx *c, int k)
{
    struct acgsolvercuda *cg = c->cg;
    struct priv *pv = c->pv;
    struct acgb200_devstate *st = pv->d_st;
    int *errcode = c->errcode;
    const int s = k & 1, n = pv->nowned;
    const int peer = c->multi && c->p2p;
    /* One allreduce for {gamma,delta} (acg/cgcuda.c:1697).  The reference
     * issues it on the compute stream ahead of q = A w, i.e. serialised.  Here:
     * peer-memory mode -- the partials were pushed to every rank right after
     * the previous update and are summed by the update kernel itself; NCCL mode
     * -- the allreduce runs on its own stream and communicator and is only
     * joined before the update.  Either way it overlaps the SpMV, which is the
     * point of pipelined CG.  {gamma_0,delta_0} were reduced during setup. */
    const int side = c->multi && !peer && pv->have_redcomm;
    if (k > 0 && c->multi && !peer) {
        if (side) {
            CU(cudaEventRecord(pv->ev_ready, pv->stream));
            CU(cudaStreamWaitEvent(pv->redstream, pv->ev_ready, 0));
            OK(acgcomm_allreduce(&st->gd_loc[s][0], &st->gd[s][0], 2, ACG_DOUBLE, ACG_SUM, pv->redstream, &pv->redcomm, errcode));
            CU(cudaEventRecord(pv->ev_red, pv->redstream));
        } else {
            OK(allreduce(c, &st->gd_loc[s][0], &st->gd[s][0], 2));
        }
    }
    OK(apply_A(c, cg->d_w, cg->d_w, cg->d_q, NULL, SPMV_Y_AX, NULL, 1, 2, 0, -1));
    if (k > 0 && side) CU(cudaStreamWaitEvent(pv->stream, pv->ev_red, 0));
    prof_mark(c, &pv->blas);
    KL(acgb200_pcg_update(n, st, 1, 0, c->multi, peer ? pv->p2p.d_desc : NULL,
                          cg->d_q, cg->d_z, cg->d_w, cg->d_t, cg->d_p, cg->d_r, c->d_x, pv->stream));
    prof_mark(c, &pv->blas);
    /* in peer-memory mode the update kernel itself pushed w's border entries
     * and {gamma,delta} of the next iteration to the other ranks, unless fusion
     * is switched off */
    if (peer && !pv->p2p.h_desc.fuse) OK(post(c, 0, -1, cg->d_w, 0, &st->gd_loc[0][0], 2, 2, 0, 0));
    c->launches += 1 + (c->multi && !peer ? 1 : 0);
    return ACG_SUCCESS;
}

int acgsolvercuda_solve_pipelined(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int tag, int *errcode,
    cublasHandle_t cublas, cusparseHandle_t cusparse)
{
    (void) cublas; (void) cusparse;
    int errcode_ = 0;
    if (!errcode) errcode = &errcode_;
    cfg_load();
    struct solvectx c;
    OK(solve_begin(&c, cg, A, b, x, diffatol, diffrtol, comm, tag, errcode));
    struct priv *pv = c.pv;
    struct acgb200_devstate *st = pv->d_st;
    const int n = pv->nowned;
    const size_t obytes = (size_t) n * sizeof(double);
    struct acgb200_devstate h;
    /* extra vectors are created on first use (acg/cgcuda.c:1167-1184) */
    OK(ensure_vec(&cg->w, &cg->d_w, A, pv->nvec, errcode));
    OK(ensure_vec(&cg->q, &cg->d_q, A, pv->nvec, errcode));
    OK(ensure_vec(&cg->z, &cg->d_z, A, pv->nvec, errcode));

    c.p2p = c.multi && pv->p2p.enabled && cfg.p2p;
    if (c.p2p) OK(acgcomm_barrier(pv->stream, comm, errcode));     /* as in acgsolvercuda_solvempi */
    if (warmup > 0) {
        memset(&h, 0, sizeof(h));
        h.maxits = warmup;
        for (int s = 0; s < 2; s++) { h.gd_loc[s][0] = h.gd[s][0] = 1; h.gd_loc[s][1] = h.gd[s][1] = 1; h.prev[s][0] = h.prev[s][1] = INFINITY; }
        OK(push_state(&c, &h));
        double *xsave = c.d_x; c.d_x = cg->d_r;
        if (c.p2p) {
            OK(acgb200_p2p_begin(&pv->p2p, warmup, pv->stream));
            OK(post(&c, 0, 0, cg->d_w, -1, NULL, 0, 0, 0, 0));
        }
        for (int i = 0; i < warmup; i++) OK(pipelined_iteration(&c, i));
        c.d_x = xsave;
        KL(acgb200_dot(n, c.d_b, c.d_b, &st->tmp_loc[0], pv->stream));
        KL(acgb200_dot2(n, cg->d_r, cg->d_w, &st->tmp_loc[0], pv->stream));
        CU(cudaStreamSynchronize(pv->stream));
    }
    if (cfg.profile) {
        OK(evpool_reserve(&pv->gemv, 2 * (maxits + 3)));
        OK(evpool_reserve(&pv->blas, 2 * (maxits + 3)));
    }
    pv->gemv.n = pv->blas.n = 0;
    /* z = t = p = 0 (acg/cgcuda.c:1516-1522) */
    CU(cudaMemsetAsync(cg->d_z, 0, obytes, pv->stream));
    CU(cudaMemsetAsync(cg->d_t, 0, obytes, pv->stream));
    CU(cudaMemsetAsync(cg->d_p, 0, obytes, pv->stream));
    CU(cudaMemsetAsync(cg->d_w, 0, ((size_t) pv->nvec + 2) * sizeof(double), pv->stream));

    cg->nsolves++; cg->niterations = 0;
    cg->bnrm2 = cg->r0nrm2 = cg->rnrm2 = cg->x0nrm2 = cg->dxnrm2 = INFINITY;
    cg->maxits = maxits; cg->diffatol = diffatol; cg->diffrtol = diffrtol;
    cg->residualatol = residualatol; cg->residualrtol = residualrtol;
    OK(acgcomm_barrier(pv->stream, comm, errcode));
    CU(cudaStreamSynchronize(pv->stream));
    const double t0 = wall();
    CU(cudaEventRecord(pv->ev_t0, pv->stream));
    c.launches = 0;

    memset(&h, 0, sizeof(h));
    OK(push_state(&c, &h));
    double bb = 0, gd0[2] = { 0, 0 };
    KL(acgb200_dot(n, c.d_b, c.d_b, &st->tmp_loc[0], pv->stream));
    OK(reduce_to_host(&c, &st->tmp_loc[0], &st->tmp[0], 1, &bb));
    cg->bnrm2 = sqrt(bb);

    /* r0 = b - A x0 ; w0 = A r0 (acg/cgcuda.c:1577-1671) */
    OK(apply_A(&c, c.d_x, c.d_x, cg->d_r, c.d_b, SPMV_R_B_AX, NULL, 0, 0, 0, -1));
    OK(apply_A(&c, cg->d_r, cg->d_r, cg->d_w, NULL, SPMV_Y_AX, NULL, 0, 0, 0, -1));
    /* gamma0 = (r0,r0), delta0 = (w0,r0): in the reference these are the first
     * two dots of the loop (acg/cgcuda.c:1680-1697); later ones come fused out
     * of the update kernel */
    KL(acgb200_dot2(n, cg->d_r, cg->d_w, &st->gd_loc[0][0], pv->stream));
    OK(reduce_to_host(&c, &st->gd_loc[0][0], &st->gd[0][0], 2, gd0));
    int converged = 0;
    double rtol_scaled = residualrtol;
    if (maxits > 0) {
        cg->rnrm2 = cg->r0nrm2 = sqrt(gd0[0]);                     /* acg/cgcuda.c:1760-1761 */
        rtol_scaled = residualrtol * cg->r0nrm2;
        const double tol = threshold(residualatol, residualrtol, cg->r0nrm2);
        memset(&h, 0, sizeof(h));
        h.maxits = maxits; h.tol = tol;
        h.gd_loc[0][0] = h.gd[0][0] = gd0[0];
        h.gd_loc[0][1] = h.gd[0][1] = gd0[1];
        h.prev[0][0] = h.prev[0][1] = INFINITY;                    /* acg/cgcuda.c:1513-1514 */
        OK(push_state(&c, &h));
        if (c.p2p) {
            /* w_0 goes to the neighbours' windows as exchange number 0 */
            OK(acgb200_p2p_begin(&pv->p2p, maxits, pv->stream));
            OK(post(&c, 0, 0, cg->d_w, -1, NULL, 0, 0, 0, 0));
        }
        OK(iterate(&c, maxits, tol > 0, 1, pipelined_iteration));
        OK(pull_state(&c, &h));
        cg->niterations = h.ctrl[0].iter;
        converged = h.ctrl[0].done;
        /* the reference reports sqrt(gamma) of the last *tested* iterate
         * (acg/cgcuda.c:1760): gamma_k at convergence, gamma_{maxits-1} otherwise */
        const int kk = converged ? cg->niterations : (cg->niterations > 0 ? cg->niterations - 1 : 0);
        const double g = converged ? h.final_rr : (kk == 0 ? gd0[0] : (c.multi ? h.gd[kk & 1][0] : h.gd_loc[kk & 1][0]));
        cg->rnrm2 = sqrt(g);
        cg->ntotaliterations += cg->niterations;
    }
    CU(cudaEventRecord(pv->ev_t1, pv->stream));
    CU(cudaEventSynchronize(pv->ev_t1));
    const double t1 = wall();
    cg->tsolve += t1 - t0;
    { float ms = 0; if (cudaEventElapsedTime(&ms, pv->ev_t0, pv->ev_t1) == cudaSuccess) pv->last_solve_ms = ms; }
    {
        const int64_t nits = cg->niterations, nnz = pv->fnnz + pv->onnz;
        const int64_t bgemv = nnz * 12 + (int64_t) n * 16 + (int64_t) (pv->nborder + pv->nghost) * 8 + (int64_t) pv->nvec * 8;
        cg->ngemv += nits + 2; cg->nflops += (nits + 2) * 3 * nnz; cg->Bgemv += (nits + 2) * bgemv;
        cg->nnrm2 += nits + 1; cg->ndot += nits; cg->nflops += nits * 4 * n; cg->Bnrm2 += nits * 8 * n; cg->Bdot += nits * 16 * n;
        cg->naxpy += nits; cg->nflops += nits * 12 * n; cg->Baxpy += nits * 56 * n;   /* acg/cgcuda.c:1783-1784 */
        if (c.multi) {
            cg->nallreduce += nits; cg->Ballreduce += nits * 16;
            cg->nhalo += nits + 2; cg->Bhalo += (nits + 2) * cg->halo->sendsize * 8;
            cg->nhalomsgs += (nits + 2) * cg->halo->nrecipients;
        }
    }
    pv->last_launches = c.launches;
    if (cfg.profile) {
        pv->last_spmv_ms = evpool_sum_ms(&pv->gemv); pv->last_spmv_n = pv->gemv.n / 2;
        pv->last_blas_ms = evpool_sum_ms(&pv->blas);
        cg->tgemv += 1e-3 * pv->last_spmv_ms;
        cg->taxpy += 1e-3 * pv->last_blas_ms;
    }
    int status = ACG_SUCCESS;
    if (!converged && !(residualatol == 0 && rtol_scaled == 0)) status = ACG_ERR_NOT_CONVERGED;
    return solve_end(&c, x, status);
}

/* ------------------------------------------------------------------------ */
/* device-resident variants                                                  */
/* ------------------------------------------------------------------------ */

/*
 * acg/cg-kernels-cuda.cu:998 / :1713.  In the reference these are the solvers
 * whose iteration never returns to the host: one cooperative kernel per solve,
 * scalars and the convergence test on the device, halo values and reductions
 * moved by device-initiated NVSHMEM operations; they require an NVSHMEM
 * communicator and a build with NVSHMEM (ACG_ERR_NVSHMEM_NOT_SUPPORTED
 * otherwise, :1012, :1727).
 *
 * In this library those three properties already hold for the loops behind
 * acgsolvercuda_solvempi / _solve_pipelined: scalars and the stopping test live
 * in the device control ring, the host only replays a captured two-iteration
 * graph and looks at the control word every few iterations, and between GPUs
 * the kernels themselves store halo values and reduction partials into the
 * peers' memory.  So with a null or NCCL communicator the device entry points
 * run those loops (same results, same report); what is *not* built is the fusion
 * of an iteration's two or three kernels into one persistent kernel with grid
 * barriers.  An NVSHMEM communicator is refused as in a
 * reference build without NVSHMEM.
 */
int acgsolvercuda_solve_device(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int *errcode)
{
    if (comm && comm->type == acgcomm_nvshmem) return ACG_ERR_NVSHMEM_NOT_SUPPORTED;
    return acgsolvercuda_solvempi(cg, A, b, x, maxits, diffatol, diffrtol, residualatol, residualrtol,
                                  warmup, comm, 0, errcode, NULL, NULL, 0);
}

int acgsolvercuda_solve_device_pipelined(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int *errcode)
{
    if (comm && comm->type == acgcomm_nvshmem) return ACG_ERR_NVSHMEM_NOT_SUPPORTED;
    return acgsolvercuda_solve_pipelined(cg, A, b, x, maxits, diffatol, diffrtol, residualatol, residualrtol,
                                         warmup, comm, 0, errcode, NULL, NULL);
}

/* ------------------------------------------------------------------------ */
/* acg/cg-kernels-cuda.h:45-97: public BLAS-1 building blocks                 */
/* ------------------------------------------------------------------------ */

static double *d_constants = NULL;      /* {-1, 1, 0}: lives as long as the process, like the reference's __constant__ symbols */

int acgsolvercuda_init_constants(double **d_minus_one, double **d_one, double **d_zero)
{
    if (!d_constants) {
        const double h[3] = { -1.0, 1.0, 0.0 };
        if (cudaMalloc((void **) &d_constants, sizeof(h)) != cudaSuccess) return ACG_ERR_CUDA;
        if (cudaMemcpy(d_constants, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) return ACG_ERR_CUDA;
    }
    *d_minus_one = d_constants; *d_one = d_constants + 1; *d_zero = d_constants + 2;
    return ACG_SUCCESS;
}

int acgsolvercuda_alpha(double *d_alpha, double *d_minus_alpha, const double *d_rnrm2sqr, const double *d_pdott)
{
    return acgb200_helper_scalars(0, d_alpha, d_minus_alpha, d_rnrm2sqr, d_pdott, 0) ? ACG_ERR_CUDA : ACG_SUCCESS;
}

int acgsolvercuda_beta(double *d_beta, const double *d_rnrm2sqr, const double *d_rnrm2sqr_prev)
{
    return acgb200_helper_scalars(1, d_beta, NULL, d_rnrm2sqr, d_rnrm2sqr_prev, 0) ? ACG_ERR_CUDA : ACG_SUCCESS;
}

int acgsolvercuda_daxpy_alpha(int n, const double *d_rnrm2sqr, const double *d_pdott, const double *d_x, double *d_y)
{
    return acgb200_helper_axpy(0, n, d_rnrm2sqr, d_pdott, d_x, d_y, 0) ? ACG_ERR_CUDA : ACG_SUCCESS;
}

int acgsolvercuda_daxpy_minus_alpha(int n, const double *d_rnrm2sqr, const double *d_pdott, const double *d_x, double *d_y)
{
    return acgb200_helper_axpy(1, n, d_rnrm2sqr, d_pdott, d_x, d_y, 0) ? ACG_ERR_CUDA : ACG_SUCCESS;
}

int acgsolvercuda_daypx_beta(int n, const double *d_rnrm2sqr, const double *d_rnrm2sqr_prev, double *d_y, const double *d_x)
{
    return acgb200_helper_axpy(2, n, d_rnrm2sqr, d_rnrm2sqr_prev, d_x, d_y, 0) ? ACG_ERR_CUDA : ACG_SUCCESS;
}

int acgsolvercuda_pipelined_daxpy_fused(
    int n, const double *d_gamma, double *d_gamma_prev, const double *d_delta, const double *d_q,
    double *d_p, double *d_r, double *d_t, double *d_x, double *d_z, double *d_w, double *d_alpha_prev,
    cudaStream_t stream)
{
    return acgb200_helper_pipelined(n, d_gamma, d_gamma_prev, d_delta, d_q, d_p, d_r, d_t, d_x, d_z, d_w, d_alpha_prev, stream)
        ? ACG_ERR_CUDA : ACG_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* report                                                                    */
/* ------------------------------------------------------------------------ */

static void opline(FILE *f, int indent, const char *name, double t, int64_t n, int64_t B)
{
    fprintf(f, "%*s  %s: %.6f seconds %" PRId64 " times %" PRId64 " B %.3f GB/s\n",
            indent, "", name, t, n, B, t > 0 ? 1.0e-9 * (double) B / t : 0.0);
}

/* same keys, order and units as acg/cgcuda.c:1893-1946, so scripts that parse
 * the reference's report keep working */
int acgsolvercuda_fwrite(FILE *f, const struct acgsolvercuda *cg, int indent)
{
    const double tother = cg->tsolve - (cg->tgemv + cg->tdot + cg->tnrm2 + cg->taxpy + cg->tcopy + cg->tallreduce + cg->thalo);
    fprintf(f, "%*sunknowns: %" PRIdx "\n", indent, "", cg->p.size);
    fprintf(f, "%*ssolves: %d\n", indent, "", cg->nsolves);
    fprintf(f, "%*stotal iterations: %d\n", indent, "", cg->ntotaliterations);
    fprintf(f, "%*stotal flops: %.3f Gflop\n", indent, "", 1.0e-9 * (double) cg->nflops);
    fprintf(f, "%*stotal flop rate: %.3f Gflop/s\n", indent, "", cg->tsolve > 0 ? 1.0e-9 * (double) cg->nflops / cg->tsolve : 0.0);
    fprintf(f, "%*stotal solver time: %.6f seconds\n", indent, "", cg->tsolve);
    fprintf(f, "%*sperformance breakdown:\n", indent, "");
    opline(f, indent, "gemv", cg->tgemv, cg->ngemv, cg->Bgemv);
    opline(f, indent, "dot", cg->tdot, cg->ndot, cg->Bdot);
    opline(f, indent, "nrm2", cg->tnrm2, cg->nnrm2, cg->Bnrm2);
    opline(f, indent, "axpy", cg->taxpy, cg->naxpy, cg->Baxpy);
    opline(f, indent, "copy", cg->tcopy, cg->ncopy, cg->Bcopy);
    opline(f, indent, "MPI_Allreduce", cg->tallreduce, cg->nallreduce, cg->Ballreduce);
    opline(f, indent, "MPI_HaloExchange", cg->thalo, cg->nhalo, cg->Bhalo);
    fprintf(f, "%*s  other: %.6f seconds\n", indent, "", tother);
    fprintf(f, "%*slast solve:\n", indent, "");
    fprintf(f, "%*s  stopping criterion:\n", indent, "");
    fprintf(f, "%*s    maximum iterations: %d\n", indent, "", cg->maxits);
    fprintf(f, "%*s    tolerance for residual: %.*g\n", indent, "", DBL_DIG, cg->residualatol);
    fprintf(f, "%*s    tolerance for relative residual: %.*g\n", indent, "", DBL_DIG, cg->residualrtol);
    fprintf(f, "%*s    tolerance for difference in solution iterates: %.*g\n", indent, "", DBL_DIG, cg->diffatol);
    fprintf(f, "%*s    tolerance for relative difference in solution iterates: %.*g\n", indent, "", DBL_DIG, cg->diffrtol);
    fprintf(f, "%*s  iterations: %d\n", indent, "", cg->niterations);
    fprintf(f, "%*s  right-hand side 2-norm: %.*g\n", indent, "", DBL_DIG, cg->bnrm2);
    fprintf(f, "%*s  initial guess 2-norm: %.*g\n", indent, "", DBL_DIG, cg->x0nrm2);
    fprintf(f, "%*s  initial residual 2-norm: %.*g\n", indent, "", DBL_DIG, cg->r0nrm2);
    fprintf(f, "%*s  residual 2-norm: %.*g\n", indent, "", DBL_DIG, cg->rnrm2);
    fprintf(f, "%*s  difference in solution iterates 2-norm: %.*g\n", indent, "", DBL_DIG, cg->dxnrm2);
    return ACG_SUCCESS;
}

#ifdef ACG_HAVE_MPI
/* acg/cgcuda.h:293 (acg/cgcuda.c:1948-2216): the same report on `root`, with
 * times taken as the maximum and flop/byte/message counters as the sum over the
 * ranks of `comm` -- the aggregation the reference performs with
 * MPI_Reduce (:1990-2012).  `verbose` adds one halo line per rank. */
int acgsolvercuda_fwritempi(FILE *f, const struct acgsolvercuda *cg, int indent, int verbose, MPI_Comm comm, int root)
{
    int rank = 0, size = 1;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    struct acgsolvercuda agg = *cg;
    double tin[8] = { cg->tsolve, cg->tgemv, cg->tdot, cg->tnrm2, cg->taxpy, cg->tcopy, cg->tallreduce, cg->thalo }, tout[8];
    int64_t cin[16] = { cg->nflops, cg->Bgemv, cg->Bdot, cg->Bnrm2, cg->Baxpy, cg->Bcopy, cg->Ballreduce, cg->Bhalo,
                        cg->nhalomsgs, 0, 0, 0, 0, 0, 0, 0 }, cout[16];
    memcpy(tout, tin, sizeof(tin)); memcpy(cout, cin, sizeof(cin));
    if (MPI_Reduce(tin, tout, 8, MPI_DOUBLE, MPI_MAX, root, comm)) return ACG_ERR_MPI;
    if (MPI_Reduce(cin, cout, 16, MPI_INT64_T, MPI_SUM, root, comm)) return ACG_ERR_MPI;
    if (rank == root) {
        agg.tsolve = tout[0]; agg.tgemv = tout[1]; agg.tdot = tout[2]; agg.tnrm2 = tout[3];
        agg.taxpy = tout[4]; agg.tcopy = tout[5]; agg.tallreduce = tout[6]; agg.thalo = tout[7];
        agg.nflops = cout[0]; agg.Bgemv = cout[1]; agg.Bdot = cout[2]; agg.Bnrm2 = cout[3]; agg.Baxpy = cout[4];
        agg.Bcopy = cout[5]; agg.Ballreduce = cout[6]; agg.Bhalo = cout[7]; agg.nhalomsgs = cout[8];
        int err = acgsolvercuda_fwrite(f, &agg, indent);
        if (err) return err;
        fprintf(f, "%*sprocesses: %d\n", indent, "", size);
    }
    if (verbose > 0 && cg->halo) {
        /* per-rank halo volume (the reference prints these under -v, :2060-2100) */
        int64_t mine[4] = { cg->halo->nrecipients, cg->halo->sendsize, cg->halo->nsenders, cg->halo->recvsize };
        int64_t *all = rank == root ? malloc((size_t) size * sizeof(mine)) : NULL;
        if (MPI_Gather(mine, 4, MPI_INT64_T, all, 4, MPI_INT64_T, root, comm)) { free(all); return ACG_ERR_MPI; }
        if (rank == root) {
            fprintf(f, "%*shalo exchange pattern (rank: recipients sendsize senders recvsize):\n", indent, "");
            for (int r = 0; r < size; r++)
                fprintf(f, "%*s  %d: %" PRId64 " %" PRId64 " %" PRId64 " %" PRId64 "\n", indent, "", r,
                        all[4 * r], all[4 * r + 1], all[4 * r + 2], all[4 * r + 3]);
            free(all);
        }
    }
    return ACG_SUCCESS;
}
#endif

/* ------------------------------------------------------------------------ */
/* extensions (include/acgb200/ext.h)                                        */
/* ------------------------------------------------------------------------ */

int acgsolvercuda_spmv(struct acgsolvercuda *cg, const double *x, double *y, int nrep, double *ms_per_spmv)
{
    int errcode_ = 0, *errcode = &errcode_;
    struct priv *pv = priv_of(cg);
    if (!pv) return ACG_ERR_INVALID_VALUE;
    if (pv->nghost > 0) return ACG_ERR_NOT_SUPPORTED;     /* whole-matrix use only */
    const int n = pv->nowned;
    CU(cudaMemcpyAsync(cg->d_p, x, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, pv->stream));
    struct acgb200_spmvargs a;
    memset(&a, 0, sizeof(a));
    a.plan = &pv->plan; a.rowptr = cg->d_rowptr; a.colidx = cg->d_colidx; a.a = cg->d_a;
    a.x = cg->d_p; a.y = cg->d_t; a.mode = SPMV_Y_AX; a.dotrows = n;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    KL(acgb200_spmv_launch(&a, pv->stream));               /* untimed first pass */
    CU(cudaEventRecord(e0, pv->stream));
    for (int i = 0; i < nrep; i++) KL(acgb200_spmv_launch(&a, pv->stream));
    CU(cudaEventRecord(e1, pv->stream));
    CU(cudaMemcpyAsync(y, cg->d_t, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, pv->stream));
    CU(cudaStreamSynchronize(pv->stream));
    float ms = 0;
    if (nrep > 0) CU(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_spmv) *ms_per_spmv = nrep > 0 ? (double) ms / nrep : 0.0;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ACG_SUCCESS;
}

/*
 * One part's share of y = A x on ONE device, ghost values supplied by the caller: x has the
 * part's nvec = owned + ghost entries, y its owned rows.  path 0 is what the set-up products and
 * the NCCL loop back-end run (local block through the tile kernel, then offdiag_kernel adding
 * the border x ghost block from the ghost tail of x: acg/cgcuda.c:858 + :878); path 1 is the
 * kernel of the peer-memory loop: border x ghost block inside the tile kernel, ghost values read
 * from a window whose senders' flags are waited for (here a loop-back window in this device's
 * memory, filled from x's tail, the flags already published).  *dot = sum_owned x_i y_i from the
 * fused epilogue.  Needs no communicator -- it exists so that the border x ghost path of a
 * partitioned matrix can be checked against the global product on a single GPU.
 */
int acgsolvercuda_spmv_ghost(struct acgsolvercuda *cg, const double *x, double *y, int path, double *dot)
{
    int errcode_ = 0, *errcode = &errcode_;
    struct priv *pv = priv_of(cg);
    if (!pv || (path != 0 && path != 1)) return ACG_ERR_INVALID_VALUE;
    const int n = pv->nowned;
    struct acgb200_devstate *st = pv->d_st;
    CU(cudaMemcpyAsync(cg->d_p, x, (size_t) pv->nvec * sizeof(double), cudaMemcpyHostToDevice, pv->stream));
    CU(cudaMemsetAsync(&st->tmp_loc[0], 0, sizeof(double), pv->stream));
    struct acgb200_spmvargs a;
    memset(&a, 0, sizeof(a));
    a.plan = &pv->plan; a.rowptr = cg->d_rowptr; a.colidx = cg->d_colidx; a.a = cg->d_a;
    a.x = cg->d_p; a.y = cg->d_t; a.mode = SPMV_Y_AX_DOT; a.acc = &st->tmp_loc[0]; a.dotrows = n; a.pub_ch = -1;
    void *d_win = NULL;
    if (path == 0) {
        a.dotrows = pv->borderoff;                  /* the border rows' share of the dot comes from offdiag_kernel */
        KL(acgb200_spmv_launch(&a, pv->stream));
        struct acgb200_offdiagargs o;
        memset(&o, 0, sizeof(o));
        o.nrows = pv->nborder; o.rowoffset = pv->borderoff;
        o.orowptr = cg->d_orowptr; o.ocolidx = cg->d_ocolidx; o.oa = cg->d_oa;
        o.x = cg->d_p; o.y = cg->d_t; o.acc = &st->tmp_loc[0]; o.dotkind = 1; o.st = st; o.p2p_iter_override = -1;
        KL(acgb200_offdiag_launch(&o, pv->stream));
    } else {
        /* loop-back window: [descriptor | flags (one per sender slot) | ghost values] */
        const size_t goff = sizeof(struct acgb200_p2pdev) + ACGB200_MAXR * sizeof(unsigned long long);
        const size_t bytes = goff + ((size_t) pv->nghost + 2) * sizeof(double);
        struct acgb200_p2pdev *h = calloc(1, sizeof(*h));
        if (!h) return ACG_ERR_ERRNO;
        cudaError_t e = cudaMalloc(&d_win, bytes);
        if (!e) e = cudaMemsetAsync(d_win, 0, bytes, pv->stream);
        if (e) { free(h); cudaFree(d_win); *errcode = (int) e; return ACG_ERR_CUDA; }
        unsigned long long *d_flags = (unsigned long long *) ((char *) d_win + sizeof(*h));
        double *d_ghost = (double *) ((char *) d_win + goff);
        h->nranks = 1; h->rank = 0;
        h->nsenders = 3; h->senders[0] = 0; h->senders[1] = 5; h->senders[2] = ACGB200_MAXR - 1;
        h->my_hflag = d_flags; h->my_ghost[0] = h->my_ghost[1] = d_ghost;
        h->hbase = 41; h->timeout_ns = 2000000000ull;
        h->borderoff = pv->borderoff; h->nborder = pv->nborder;
        unsigned long long flags[ACGB200_MAXR];
        for (int i = 0; i < ACGB200_MAXR; i++) flags[i] = 0;
        flags[0] = 41; flags[5] = 42; flags[ACGB200_MAXR - 1] = 41;      /* published: sequence >= hbase + iteration 0 */
        e = cudaMemcpyAsync(d_win, h, sizeof(*h), cudaMemcpyHostToDevice, pv->stream);
        if (!e) e = cudaMemcpyAsync(d_flags, flags, sizeof(flags), cudaMemcpyHostToDevice, pv->stream);
        if (!e && pv->nghost > 0)
            e = cudaMemcpyAsync(d_ghost, cg->d_p + n, (size_t) pv->nghost * sizeof(double), cudaMemcpyDeviceToDevice, pv->stream);
        /* the tail of x itself is poisoned: path 1 must take the ghosts from the window */
        if (!e && pv->nghost > 0) e = cudaMemsetAsync(cg->d_p + n, 0xff, (size_t) pv->nghost * sizeof(double), pv->stream);
        if (!e) e = cudaStreamSynchronize(pv->stream);
        free(h);
        if (e) { cudaFree(d_win); *errcode = (int) e; return ACG_ERR_CUDA; }
        a.p2p = (const struct acgb200_p2pdev *) d_win;
        a.od_rowoffset = pv->borderoff; a.od_nrows = pv->nborder;
        a.orowptr = cg->d_orowptr; a.ocolidx = cg->d_ocolidx; a.oa = cg->d_oa;
        int le = acgb200_spmv_launch(&a, pv->stream);
        if (le) { cudaFree(d_win); *errcode = le; return ACG_ERR_CUDA; }
    }
    double hdot = 0;
    cudaError_t e = cudaMemcpyAsync(y, cg->d_t, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, pv->stream);
    if (!e) e = cudaMemcpyAsync(&hdot, &st->tmp_loc[0], sizeof(double), cudaMemcpyDeviceToHost, pv->stream);
    if (!e) e = cudaStreamSynchronize(pv->stream);
    cudaFree(d_win);
    CU(e);
    if (dot) *dot = hdot;
    return ACG_SUCCESS;
}

int acgsolvercuda_info(const struct acgsolvercuda *cg, struct acgb200_info *info)
{
    const struct priv *pv = priv_of(cg);
    if (!pv) return ACG_ERR_INVALID_VALUE;
    memset(info, 0, sizeof(*info));
    info->spmv_lanes_per_row = pv->plan.lanes_per_row;
    info->spmv_rows_cap = pv->plan.rows_cap; info->spmv_nnz_cap = pv->plan.nnz_cap;
    info->spmv_stages = pv->plan.nstages; info->spmv_ntiles = pv->plan.ntiles;
    info->spmv_nlong = pv->plan.nlong; info->spmv_grid = pv->plan.grid; info->spmv_smem_bytes = pv->plan.smem_bytes;
    info->last_launches = pv->last_launches;
    info->last_spmv_ms = pv->last_spmv_ms; info->last_spmv_count = pv->last_spmv_n;
    info->last_solve_ms = pv->last_solve_ms;
    info->last_h2d_ms = pv->last_h2d_ms; info->last_d2h_ms = pv->last_d2h_ms;
    info->last_blas_ms = pv->last_blas_ms;
    info->num_sms = acgb200_num_sms();
    info->spmv_merge_tiles = pv->plan.nmtiles; info->spmv_merge_rows = pv->plan.merge_rows; info->spmv_merge_split = pv->plan.nsplit;
    info->spmv_slices = pv->plan.nslices; info->spmv_slice_rows = pv->plan.slice_rows;
    info->spmv_slice_exc = (int) pv->plan.slice_exc;
    info->spmv_slice_ub = pv->plan.slice_ub; info->spmv_slice_grid = pv->plan.slice_grid;
    info->spmv_nmedium = pv->plan.nmed;
    info->spmv_min_bytes = acgb200_spmv_min_bytes(&pv->plan);
    return ACG_SUCCESS;
}
