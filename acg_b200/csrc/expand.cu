/*
 * expand.cu -- full-storage expansion of the packed symmetric CSR matrix on the
 * device: the GPU form of acgsymcsrmatrix_dsymv_init (acg/symcsrmatrix.c:760-851,
 * SURVEY.md 8(f) item 2).
 *
 * The reference mirrors the packed triangle with a serial scatter loop over the
 * rows i = 0..n-1: entry (i,j) goes to the next free slot of row i and, when
 * i != j, of row j.  The resulting order inside a full row j is therefore
 *
 *     [ entries (i,j) of rows i < j, ascending i ]      "low" transposed entries
 *     [ row j's own packed entries, in packed order ]
 *     [ entries (i,j) of rows i > j, ascending i ]      (local renumbering after a
 *                                                        partition makes these possible)
 *
 * and the SpMV adds a row's products in exactly this order, so the order is part
 * of the result.  Here it is reproduced without any serial pass:
 *
 *   1. expand_count_kernel   a warp per packed row: counts per row (own entries,
 *                            border x ghost entries) and, with atomics, per target
 *                            row (transposed entries, and how many of them come from
 *                            smaller rows); emits one sort record per packed entry:
 *                            key = target row j (or the sentinel n), value = (i, k).
 *   2. exclusive scans       full row pointers, transposed-segment starts, border x
 *                            ghost row pointers (cub::DeviceScan).
 *   3. stable radix sort     of the records by key (cub::DeviceRadixSort): records
 *                            are emitted in (i,k) order, stability keeps it inside a
 *                            key -- exactly the reference's "ascending i" order.
 *   4. expand_fill_own       a warp per packed row copies the row's own entries (order
 *                            preserved by ballot/popc compaction) behind the low
 *                            transposed ones, and the ghost-column entries into the
 *                            border x ghost block (columns rebased by -borderrowoffset).
 *   5. expand_fill_transposed   a thread per sorted record places (j,i,a[k]).
 *
 * Output arrays are byte-identical to the host fill (tests/test_gpu_expand.py) and
 * carry the padding the TMA tile copies need (internal.h).  The sort and the scans
 * are library code (CUB, header-only, compiled into this object); they run once
 * per matrix, outside the CG loop.
 */
#include "internal.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define EXP_THREADS 256

__device__ __forceinline__ int warp_sum_int(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* one warp per packed row */
__global__ void __launch_bounds__(EXP_THREADS)
expand_count_kernel(int n, int ghost0, int border0, const int *__restrict__ rp, const int *__restrict__ col,
                    int *__restrict__ own, int *__restrict__ low, int *__restrict__ tcnt, int *__restrict__ ocnt,
                    int *__restrict__ keys, unsigned long long *__restrict__ vals)
{
    const int lane = threadIdx.x & 31;
    const int warp = (int) (((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int) (((size_t) gridDim.x * blockDim.x) >> 5);
    for (int i = warp; i < n; i += nwarps) {
        const int kb = rp[i], ke = rp[i + 1];
        int nown = 0, noff = 0;
        for (int k = kb + lane; k < ke; k += 32) {
            const int j = col[k];
            int key = n;                          /* sentinel: sorts behind every row */
            if (j < ghost0) {
                nown++;
                if (j != i) {
                    key = j;
                    atomicAdd(&tcnt[j], 1);
                    if (i < j) atomicAdd(&low[j], 1);
                }
            } else {
                noff++;
            }
            keys[k] = key;
            vals[k] = ((unsigned long long) (unsigned) i << 32) | (unsigned) k;
        }
        __syncwarp();
        nown = warp_sum_int(nown);
        noff = warp_sum_int(noff);
        if (lane == 0) {
            own[i] = nown;
            if (i >= border0) ocnt[i - border0] = noff;
        }
    }
}

/* tot[i] = own[i] + tcnt[i] (i < n), tot[n] = 0: input of the row-pointer scan */
__global__ void __launch_bounds__(EXP_THREADS)
expand_total_kernel(int n, const int *__restrict__ own, const int *__restrict__ tcnt, int *__restrict__ tot)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x)
        tot[i] = i < n ? own[i] + tcnt[i] : 0;
}

/* one warp per packed row: own entries and border x ghost entries, packed order preserved */
__global__ void __launch_bounds__(EXP_THREADS)
expand_fill_own_kernel(int n, int ghost0, int border0, double eps,
                       const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ a,
                       const int *__restrict__ frp, const int *__restrict__ low, const int *__restrict__ orp,
                       int *__restrict__ fcol, double *__restrict__ fa, int *__restrict__ ocol, double *__restrict__ oa)
{
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const int warp = (int) (((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int) (((size_t) gridDim.x * blockDim.x) >> 5);
    for (int i = warp; i < n; i += nwarps) {
        const int kb = rp[i], ke = rp[i + 1];
        const int base_f = frp[i] + low[i];
        const bool has_o = i >= border0;
        const int base_o = has_o ? orp[i - border0] : 0;
        int cf = 0, co = 0;
        for (int k0 = kb; k0 < ke; k0 += 32) {          /* warp-uniform trip count */
            const int k = k0 + lane;
            const bool valid = k < ke;
            const int j = valid ? col[k] : 0;
            const double av = valid ? a[k] : 0.0;
            const bool isf = valid && j < ghost0;
            const bool iso = valid && j >= ghost0;
            const unsigned mf = __ballot_sync(0xffffffffu, isf);
            const unsigned mo = __ballot_sync(0xffffffffu, iso);
            if (isf) {
                const int pos = base_f + cf + __popc(mf & below);
                fcol[pos] = j;
                fa[pos] = av + (j == i ? eps : 0.0);
            }
            if (iso && has_o) {
                const int pos = base_o + co + __popc(mo & below);
                ocol[pos] = j - border0;
                oa[pos] = av;
            }
            cf += __popc(mf);
            co += __popc(mo);
        }
    }
}

/* one thread per sorted record */
__global__ void __launch_bounds__(EXP_THREADS)
expand_fill_transposed_kernel(int64_t nrec, int n, const int *__restrict__ skeys, const unsigned long long *__restrict__ svals,
                              const double *__restrict__ a, const int *__restrict__ frp, const int *__restrict__ tstart,
                              const int *__restrict__ own, int *__restrict__ fcol, double *__restrict__ fa)
{
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < nrec; s += (int64_t) gridDim.x * blockDim.x) {
        const int j = skeys[s];
        if (j >= n) continue;                           /* diagonal / ghost-column entries: no transposed copy */
        const unsigned long long v = svals[s];
        const int i = (int) (v >> 32);
        const unsigned k = (unsigned) (v & 0xffffffffull);
        const int r = (int) (s - tstart[j]);
        const int pos = frp[j] + r + (i > j ? own[j] : 0);
        fcol[pos] = i;
        fa[pos] = a[k];
    }
}

/* pad[i] = last for i in [0,npad) */
__global__ void expand_pad_kernel(int *p, int npad, const int *last)
{
    const int v = *last;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) p[i] = v;
}

static int grid_for(int64_t work, int per_block)
{
    int64_t g = (work + per_block - 1) / per_block;
    const int64_t cap = (int64_t) acgb200_num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int) g;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

/*
 * Inputs: the packed matrix on the device (row pointers int32 [n+1], 0-based
 * columns and values [pnnz]).  n = nprows (owned + ghost rows), nob = nborderrows
 * + nghostrows.  Outputs (cudaMalloc'ed here, owned by the caller afterwards):
 * full local block (row pointers [n+1+rp_pad], columns/values [fnnz+blk_pad])
 * and border x ghost block ([nob+1+rp_pad], [onnz+blk_pad]).  Returns a
 * cudaError_t value (0 = success); cudaErrorInvalidValue if the full storage
 * would not fit 32-bit offsets.
 */
extern "C" int acgb200_expand_device(int n, int ghost0, int border0, int nob, int64_t pnnz,
                                     const int *d_prp, const int *d_pcol, const double *d_pa, double eps,
                                     int rp_pad, int blk_pad, struct acgb200_expanded *out, cudaStream_t stream)
{
    cudaError_t err = cudaSuccess;
    int *own = NULL, *low = NULL, *tcnt = NULL, *ocnt = NULL, *tot = NULL, *tstart = NULL;
    int *keys = NULL, *skeys = NULL;
    unsigned long long *vals = NULL, *svals = NULL;
    void *tmp = NULL;
    size_t tmp_bytes = 0, need = 0;
    int *frp = NULL, *fcol = NULL, *orp = NULL, *ocol = NULL;
    double *fa = NULL, *oa = NULL;
    int fnnz = 0, onnz = 0;
    int end_bit = 1;
    const size_t rec = (size_t) (pnnz > 0 ? pnnz : 1);
    memset(out, 0, sizeof(*out));
    if (n < 0 || nob < 0 || pnnz < 0 || 2 * pnnz > (int64_t) INT32_MAX) return (int) cudaErrorInvalidValue;
    while (end_bit < 32 && (1ll << end_bit) <= (long long) n) end_bit++;       /* keys are in [0, n] */

    CK(cudaMalloc((void **) &own, ((size_t) n + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &low, ((size_t) n + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &tcnt, ((size_t) n + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &ocnt, ((size_t) nob + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &tot, ((size_t) n + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &tstart, ((size_t) n + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &keys, rec * sizeof(int)));
    CK(cudaMalloc((void **) &skeys, rec * sizeof(int)));
    CK(cudaMalloc((void **) &vals, rec * sizeof(unsigned long long)));
    CK(cudaMalloc((void **) &svals, rec * sizeof(unsigned long long)));
    CK(cudaMalloc((void **) &frp, ((size_t) n + 1 + (size_t) rp_pad) * sizeof(int)));
    CK(cudaMalloc((void **) &orp, ((size_t) nob + 1 + (size_t) rp_pad) * sizeof(int)));
    CK(cudaMemsetAsync(low, 0, ((size_t) n + 1) * sizeof(int), stream));
    CK(cudaMemsetAsync(tcnt, 0, ((size_t) n + 1) * sizeof(int), stream));
    CK(cudaMemsetAsync(ocnt, 0, ((size_t) nob + 1) * sizeof(int), stream));

    /* temporary storage of the library calls: the largest of the four requests */
    CK(cub::DeviceScan::ExclusiveSum(NULL, need, tot, frp, n + 1, stream));
    tmp_bytes = need;
    CK(cub::DeviceScan::ExclusiveSum(NULL, need, ocnt, orp, nob + 1, stream));
    if (need > tmp_bytes) tmp_bytes = need;
    CK(cub::DeviceRadixSort::SortPairs(NULL, need, keys, skeys, vals, svals, (int64_t) pnnz, 0, end_bit, stream));
    if (need > tmp_bytes) tmp_bytes = need;
    CK(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 16));

    /* 1. counts and sort records */
    if (n > 0) {
        expand_count_kernel<<<grid_for((int64_t) n * 32, EXP_THREADS), EXP_THREADS, 0, stream>>>(
            n, ghost0, border0, d_prp, d_pcol, own, low, tcnt, ocnt, keys, vals);
        CK(cudaGetLastError());
    }
    /* 2. scans */
    expand_total_kernel<<<grid_for((int64_t) n + 1, EXP_THREADS), EXP_THREADS, 0, stream>>>(n, own, tcnt, tot);
    CK(cudaGetLastError());
    need = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(tmp, need, tot, frp, n + 1, stream));
    need = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(tmp, need, tcnt, tstart, n + 1, stream));
    need = tmp_bytes;
    CK(cub::DeviceScan::ExclusiveSum(tmp, need, ocnt, orp, nob + 1, stream));
    CK(cudaMemcpyAsync(&fnnz, frp + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&onnz, orp + nob, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (rp_pad > 0) {
        expand_pad_kernel<<<1, 32, 0, stream>>>(frp + n + 1, rp_pad, frp + n);
        expand_pad_kernel<<<1, 32, 0, stream>>>(orp + nob + 1, rp_pad, orp + nob);
        CK(cudaGetLastError());
    }
    CK(cudaMalloc((void **) &fcol, ((size_t) fnnz + (size_t) blk_pad + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &fa, ((size_t) fnnz + (size_t) blk_pad + 1) * sizeof(double)));
    CK(cudaMalloc((void **) &ocol, ((size_t) onnz + (size_t) blk_pad + 1) * sizeof(int)));
    CK(cudaMalloc((void **) &oa, ((size_t) onnz + (size_t) blk_pad + 1) * sizeof(double)));
    CK(cudaMemsetAsync(fcol + fnnz, 0, ((size_t) blk_pad + 1) * sizeof(int), stream));
    CK(cudaMemsetAsync(fa + fnnz, 0, ((size_t) blk_pad + 1) * sizeof(double), stream));
    CK(cudaMemsetAsync(ocol + onnz, 0, ((size_t) blk_pad + 1) * sizeof(int), stream));
    CK(cudaMemsetAsync(oa + onnz, 0, ((size_t) blk_pad + 1) * sizeof(double), stream));
    /* 3. stable sort of the records by target row */
    if (pnnz > 0) {
        need = tmp_bytes;
        CK(cub::DeviceRadixSort::SortPairs(tmp, need, keys, skeys, vals, svals, (int64_t) pnnz, 0, end_bit, stream));
    }
    /* 4., 5. fill */
    if (n > 0) {
        expand_fill_own_kernel<<<grid_for((int64_t) n * 32, EXP_THREADS), EXP_THREADS, 0, stream>>>(
            n, ghost0, border0, eps, d_prp, d_pcol, d_pa, frp, low, orp, fcol, fa, ocol, oa);
        CK(cudaGetLastError());
    }
    if (pnnz > 0) {
        expand_fill_transposed_kernel<<<grid_for(pnnz, EXP_THREADS), EXP_THREADS, 0, stream>>>(
            pnnz, n, skeys, svals, d_pa, frp, tstart, own, fcol, fa);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(stream));
    out->d_rowptr = frp; out->d_colidx = fcol; out->d_a = fa;
    out->d_orowptr = orp; out->d_ocolidx = ocol; out->d_oa = oa;
    out->fnnz = fnnz; out->onnz = onnz;
    frp = NULL; fcol = NULL; fa = NULL; orp = NULL; ocol = NULL; oa = NULL;
done:
    cudaFree(own); cudaFree(low); cudaFree(tcnt); cudaFree(ocnt); cudaFree(tot); cudaFree(tstart);
    cudaFree(keys); cudaFree(skeys); cudaFree(vals); cudaFree(svals); cudaFree(tmp);
    cudaFree(frp); cudaFree(fcol); cudaFree(fa); cudaFree(orp); cudaFree(ocol); cudaFree(oa);
    return (int) err;
}
