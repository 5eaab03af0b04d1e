/*
 * acgb200/cgcuda.h -- the drop-in boundary: conjugate-gradient solver on B200.
 *
 * ABI counterpart of acg/cgcuda.h:68-117 (struct acgsolvercuda) and :126-300
 * plus acg/cg-kernels-cuda.h:124-175 (the device-resident variants).  The
 * reference driver cuda/acg-cuda.c calls exactly these (:2209 init, :2242-2262
 * solve dispatch, :2270 report, :2426 free) and links against this library
 * unchanged; see INTEGRATION.md.
 *
 * What differs behind the boundary: the two cusparseSpMV calls,
 * two cublasDdot calls and three axpy kernels per iteration of
 * acg/cgcuda.c:845-1019 become three hand-written sm_100a kernels (a
 * TMA-staged CSR SpMV with the p.Ap dot folded in, r-update with r.r folded
 * in, fused x/p update), scalars stay on the device, and the host polls for
 * convergence every few iterations instead of synchronising on every one.
 * The cuBLAS/cuSPARSE handles are accepted for signature compatibility and
 * never used.
 */
#ifndef ACGB200_CGCUDA_H
#define ACGB200_CGCUDA_H

#include "acgb200/config.h"
#include "acgb200/comm.h"
#include "acgb200/halo.h"
#include "acgb200/symcsrmatrix.h"
#include "acgb200/vector.h"

#include <stdint.h>
#include <stdio.h>

/* Opaque stand-ins for the library handle types in the reference signatures
 * (acg/cgcuda.h:140-147, :208-225); identical to the definitions in
 * cublas_v2.h / cusparse.h, which take precedence if already included. */
#ifndef CUBLAS_API_H_
typedef struct cublasContext *cublasHandle_t;
#endif
#ifndef CUSPARSE_H_
typedef struct cusparseContext *cusparseHandle_t;
typedef int cusparseSpMVAlg_t;   /* enum passed by value: int in the C ABI */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* acg/cgcuda.h:68-117 */
struct acgsolvercuda {
    struct acgvector r, p, t;
    struct acgvector *w, *q, *z;
    struct acgvector *dx;
    struct acghalo *halo;
    struct acghaloexchange *haloexchange;
    int maxits;
    double diffatol, diffrtol, residualatol, residualrtol;
    double bnrm2, r0nrm2, rnrm2, x0nrm2, dxnrm2;
    double *d_minus_one, *d_one, *d_zero, *d_inf;
    double *d_bnrm2sqr, *d_r0nrm2sqr, *d_rnrm2sqr, *d_rnrm2sqr_prev;
    double *d_pdott, *d_alpha, *d_minus_alpha, *d_beta;
    int *d_niterations, *d_converged;
    double *d_r, *d_p, *d_t, *d_w, *d_q, *d_z;
    acgidx_t *d_rowptr, *d_orowptr;
    acgidx_t *d_colidx, *d_ocolidx;
    double *d_a, *d_oa;
    int use_nvshmem;
    int nsolves, ntotaliterations, niterations;
    int64_t nflops;
    double tsolve;
    double tgemv, tdot, tnrm2, taxpy, tcopy, tallreduce, thalo;
    int64_t ngemv, ndot, nnrm2, naxpy, ncopy, nallreduce, nhalo;
    int64_t Bgemv, Bdot, Bnrm2, Baxpy, Bcopy, Ballreduce, Bhalo;
    int64_t nhalomsgs;
};

/* acg/cgcuda.h:126 */
ACG_API void acgsolvercuda_free(struct acgsolvercuda *cg);

/* acg/cgcuda.h:140 -- uploads both CSR blocks, builds the SpMV tile plan, the
 * halo pattern and all device state (acg/cgcuda.c:138-330) */
ACG_API int acgsolvercuda_init(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    cublasHandle_t cublas, cusparseHandle_t cusparse, const struct acgcomm *comm);

/* acg/cgcuda.h:208 -- classic CG, host-driven (acg/cgcuda.c:398-1108).
 * b and x are HOST vectors; they are copied to the device on entry and x is
 * copied back on exit, exactly as the reference does (:484-493, :1063). */
ACG_API int acgsolvercuda_solvempi(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int tag, int *errcode,
    cublasHandle_t cublas, cusparseHandle_t cusparse, cusparseSpMVAlg_t cusparse_spmv_alg);

/* acg/cgcuda.h:251 -- pipelined (Ghysels-Vanroose) CG (acg/cgcuda.c:1136-1880) */
ACG_API int acgsolvercuda_solve_pipelined(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int tag, int *errcode,
    cublasHandle_t cublas, cusparseHandle_t cusparse);

/* acg/cgcuda.h:167 -- single-process convenience (declared, never defined, in
 * the reference: acg/cgcuda.c:357-367); here: classic CG with a null comm */
ACG_API int acgsolvercuda_solve(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup);

/* acg/cg-kernels-cuda.h:45-97 -- the BLAS-1 building blocks of the reference's loops, operating on
 * device scalars and vectors (acg/cg-kernels-cuda.cu:54-303).  The solvers of this library do not
 * use them (their updates are fused, kernels.cu); they are kept for callers written against the
 * reference's header.  Same semantics: alpha = rr/(p,Ap), beta = rr/rr_prev read on the device,
 * launches on the legacy default stream except where a stream is passed. */
ACG_API int acgsolvercuda_init_constants(double **d_minus_one, double **d_one, double **d_zero);      /* :45 */
ACG_API int acgsolvercuda_alpha(double *d_alpha, double *d_minus_alpha, const double *d_rnrm2sqr, const double *d_pdott);   /* :50 */
ACG_API int acgsolvercuda_beta(double *d_beta, const double *d_rnrm2sqr, const double *d_rnrm2sqr_prev);                  /* :56 */
ACG_API int acgsolvercuda_daxpy_alpha(int n, const double *d_rnrm2sqr, const double *d_pdott, const double *d_x, double *d_y);        /* :61  y += (rr/pAp) x */
ACG_API int acgsolvercuda_daxpy_minus_alpha(int n, const double *d_rnrm2sqr, const double *d_pdott, const double *d_x, double *d_y);  /* :68  y -= (rr/pAp) x */
ACG_API int acgsolvercuda_pipelined_daxpy_fused(                                                       /* :76 */
    int n, const double *d_gamma, double *d_gamma_prev, const double *d_delta, const double *d_q,
    double *d_p, double *d_r, double *d_t, double *d_x, double *d_z, double *d_w, double *d_alpha_prev,
    cudaStream_t stream);
ACG_API int acgsolvercuda_daypx_beta(int n, const double *d_rnrm2sqr, const double *d_rnrm2sqr_prev, double *d_y, const double *d_x); /* :91  y = (rr/rr_prev) y + x */

/* acg/cg-kernels-cuda.h:124, :163 -- device-resident solvers.  The reference implements them on
 * NVSHMEM only (one cooperative kernel per solve) and otherwise returns
 * ACG_ERR_NVSHMEM_NOT_SUPPORTED (acg/cg-kernels-cuda.cu:1012, :1727).  Here: with a null or NCCL
 * communicator they run the classic / pipelined loop, whose control and inter-GPU communication are
 * device-resident already (cgcuda.c); an NVSHMEM communicator gets ACG_ERR_NVSHMEM_NOT_SUPPORTED. */
ACG_API int acgsolvercuda_solve_device(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int *errcode);
ACG_API int acgsolvercuda_solve_device_pipelined(
    struct acgsolvercuda *cg, const struct acgsymcsrmatrix *A,
    const struct acgvector *b, struct acgvector *x,
    int maxits, double diffatol, double diffrtol, double residualatol, double residualrtol,
    int warmup, struct acgcomm *comm, int *errcode);

/* acg/cgcuda.h:280 -- solver report (acg/cgcuda.c:1893-1946) */
ACG_API int acgsolvercuda_fwrite(FILE *f, const struct acgsolvercuda *cg, int indent);
#ifdef ACG_HAVE_MPI
/* acg/cgcuda.h:293 -- (acg/cgcuda.c:1948-2216) */
ACG_API int acgsolvercuda_fwritempi(FILE *f, const struct acgsolvercuda *cg, int indent, int verbose, MPI_Comm comm, int root);
#endif

#ifdef __cplusplus
}
#endif
#endif
