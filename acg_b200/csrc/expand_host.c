/*
 * expand_host.c -- host side of the device full-storage expansion (expand.cu):
 * the packed symmetric CSR goes to the GPU (1.9 GB at C3 instead of the 3.65 GB
 * of the expanded arrays), acgb200_expand_device builds the local block and the
 * border x ghost block there, byte-identical to acgsymcsrmatrix_dsymv_init
 * (acg/symcsrmatrix.c:760-851).
 *
 *   acgb200_expand_upload              packed arrays -> device -> expanded device arrays
 *                                      (used by acgsolvercuda_init when the matrix has no full storage)
 *   acgsymcsrmatrix_dsymv_init_cuda    the same contract as acgsymcsrmatrix_dsymv_init, computed
 *                                      on the device and copied back into the matrix
 */
#include "acgb200/error.h"
#include "acgb200/ext.h"
#include "acgb200/symcsrmatrix.h"
#include "hostmem.h"
#include "internal.h"

#include <cuda_runtime_api.h>
#include <stdlib.h>
#include <string.h>

#define CUE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (errcode) *errcode = (int) e_; err = ACG_ERR_CUDA; goto done; } } while (0)

void acgb200_expanded_free(struct acgb200_expanded *x)
{
    cudaFree(x->d_rowptr); cudaFree(x->d_colidx); cudaFree(x->d_a);
    cudaFree(x->d_orowptr); cudaFree(x->d_ocolidx); cudaFree(x->d_oa);
    memset(x, 0, sizeof(*x));
}

int acgb200_expand_upload(const struct acgsymcsrmatrix *A, double eps, int rp_pad, int blk_pad,
                          struct acgb200_expanded *out, cudaStream_t stream, int *errcode)
{
    int err = ACG_SUCCESS;
    const int64_t n = A->nprows, pnnz = A->rowptr ? A->rowptr[n] : 0;
    const int base = A->rowidxbase;
    int *h_rp = NULL, *h_col = NULL, *d_rp = NULL, *d_col = NULL;
    double *d_a = NULL;
    memset(out, 0, sizeof(*out));
    if (!A->rowptr || (pnnz > 0 && (!A->colidx || !A->a))) return ACG_ERR_INVALID_VALUE;
    if (2 * pnnz > (int64_t) INT32_MAX || n >= (int64_t) INT32_MAX) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    h_rp = malloc(((size_t) n + 1) * sizeof(*h_rp));
    if (!h_rp) return ACG_ERR_ERRNO;
    for (int64_t i = 0; i <= n; i++) h_rp[i] = (int) A->rowptr[i];
    CUE(cudaMalloc((void **) &d_rp, ((size_t) n + 1) * sizeof(int)));
    CUE(cudaMalloc((void **) &d_col, (size_t) (pnnz > 0 ? pnnz : 1) * sizeof(int)));
    CUE(cudaMalloc((void **) &d_a, (size_t) (pnnz > 0 ? pnnz : 1) * sizeof(double)));
    CUE(cudaMemcpyAsync(d_rp, h_rp, ((size_t) n + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (pnnz > 0) {
        if (base == 0) {
            CUE(cudaMemcpyAsync(d_col, A->colidx, (size_t) pnnz * sizeof(int), cudaMemcpyHostToDevice, stream));
        } else {
            h_col = malloc((size_t) pnnz * sizeof(*h_col));
            if (!h_col) { err = ACG_ERR_ERRNO; goto done; }
            for (int64_t k = 0; k < pnnz; k++) h_col[k] = A->colidx[k] - base;
            CUE(cudaMemcpyAsync(d_col, h_col, (size_t) pnnz * sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        CUE(cudaMemcpyAsync(d_a, A->a, (size_t) pnnz * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    CUE(cudaStreamSynchronize(stream));
    {
        const int ce = acgb200_expand_device((int) n, A->ghostrowoffset, A->borderrowoffset, A->nborderrows + A->nghostrows,
                                             pnnz, d_rp, d_col, d_a, eps, rp_pad, blk_pad, out, stream);
        if (ce) { if (errcode) *errcode = ce; err = ce == (int) cudaErrorInvalidValue ? ACG_ERR_INDEX_OUT_OF_BOUNDS : ACG_ERR_CUDA; }
    }
done:
    free(h_rp); free(h_col);
    cudaFree(d_rp); cudaFree(d_col); cudaFree(d_a);
    return err;
}

int acgsymcsrmatrix_dsymv_init_cuda(struct acgsymcsrmatrix *A, double eps, int *errcode)
{
    int err = ACG_SUCCESS;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return ACG_ERR_CUDA;   /* no CPU fallback: use acgsymcsrmatrix_dsymv_init */
    struct acgb200_expanded x;
    err = acgb200_expand_upload(A, eps, 0, 0, &x, 0, errcode);
    if (err) return err;
    const int64_t n = A->nprows, no = (int64_t) A->nborderrows + A->nghostrows;
    const int base = A->rowidxbase;
    int *t = NULL;
    free(A->frowptr); free(A->fcolidx); free(A->fa);
    free(A->orowptr); free(A->ocolidx); free(A->oa);
    A->frowptr = NULL; A->fcolidx = NULL; A->fa = NULL; A->orowptr = NULL; A->ocolidx = NULL; A->oa = NULL;
    A->fnpnzs = x.fnnz; A->onpnzs = x.onnz;
    A->frowptr = acgb200_bigalloc(((size_t) n + 1) * sizeof(*A->frowptr));
    A->fcolidx = acgb200_bigalloc((size_t) (x.fnnz > 0 ? x.fnnz : 1) * sizeof(*A->fcolidx));
    A->fa = acgb200_bigalloc((size_t) (x.fnnz > 0 ? x.fnnz : 1) * sizeof(*A->fa));
    A->orowptr = malloc(((size_t) no + 1) * sizeof(*A->orowptr));
    A->ocolidx = malloc((size_t) (x.onnz > 0 ? x.onnz : 1) * sizeof(*A->ocolidx));
    A->oa = malloc((size_t) (x.onnz > 0 ? x.onnz : 1) * sizeof(*A->oa));
    t = malloc(((size_t) (n > no ? n : no) + 1) * sizeof(*t));
    if (!A->frowptr || !A->fcolidx || !A->fa || !A->orowptr || !A->ocolidx || !A->oa || !t) { err = ACG_ERR_ERRNO; goto done; }
    CUE(cudaMemcpy(t, x.d_rowptr, ((size_t) n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i <= n; i++) A->frowptr[i] = t[i];
    CUE(cudaMemcpy(t, x.d_orowptr, ((size_t) no + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i <= no; i++) A->orowptr[i] = t[i];
    if (x.fnnz > 0) {
        CUE(cudaMemcpy(A->fcolidx, x.d_colidx, (size_t) x.fnnz * sizeof(int), cudaMemcpyDeviceToHost));
        CUE(cudaMemcpy(A->fa, x.d_a, (size_t) x.fnnz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (x.onnz > 0) {
        CUE(cudaMemcpy(A->ocolidx, x.d_ocolidx, (size_t) x.onnz * sizeof(int), cudaMemcpyDeviceToHost));
        CUE(cudaMemcpy(A->oa, x.d_oa, (size_t) x.onnz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (base != 0) {
        for (int64_t k = 0; k < x.fnnz; k++) A->fcolidx[k] += base;
        for (int64_t k = 0; k < x.onnz; k++) A->ocolidx[k] += base;
    }
done:
    free(t);
    acgb200_expanded_free(&x);
    if (err) {          /* never leave half-filled full storage behind: the solver would take it for valid */
        free(A->frowptr); free(A->fcolidx); free(A->fa);
        free(A->orowptr); free(A->ocolidx); free(A->oa);
        A->frowptr = NULL; A->fcolidx = NULL; A->fa = NULL; A->orowptr = NULL; A->ocolidx = NULL; A->oa = NULL;
        A->fnpnzs = 0; A->onpnzs = 0;
    }
    return err;
}
