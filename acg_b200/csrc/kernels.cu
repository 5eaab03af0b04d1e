/*
 * kernels.cu -- hand-written sm_100a kernels of the CG hot path.
 *
 * What each kernel replaces in the reference (paths relative to the aCG tree):
 *
 *   spmv_tiles_kernel     cusparseSpMV(matA)  acg/cgcuda.c:858,:1724 (+:775,:1639)
 *                         + cublasDdot(p,t)   acg/cgcuda.c:894  (fused epilogue)
 *   spmv_long_*           same, rows longer than one tile (power-law inputs);
 *                         the role csrgemv_merge plays in acg/cg-kernels-cuda.cu:340
 *   offdiag_kernel        cusparseSpMV(matO)  acg/cgcuda.c:878,:1744; csrgemv :443
 *   cg_update_r_kernel    daxpy_minus_alpha   acg/cg-kernels-cuda.cu:153 + cublasDdot(r,r) acg/cgcuda.c:933
 *   cg_update_xp_kernel   daxpy_alpha :119 + daypx_beta :271 (one pass instead of two)
 *   pcg_update_kernel     pipelined_daxpy_fused :187 + cublasDdot x2 acg/cgcuda.c:1680,:1688
 *   gather/scatter        acghalo_pack/unpack_cuda_double acg/halo.cu:41,:94
 *
 * Design (details in DESIGN.md): FP64 SIMT, HBM-bound, tensor cores unused.
 * The CSR streams (values, column indices, row pointers) are moved by the TMA
 * engine -- 1-D cp.async.bulk global->shared copies completing on an mbarrier
 * -- into a multi-stage shared-memory ring, so the SMs' load/store units only
 * issue the x gathers and the y stores.  Rows are processed from shared memory
 * by G-lane groups (G=1: one thread per row, the warp's gathers of 32
 * consecutive rows then coalesce for banded/stencil matrices).  Dot products
 * are folded into the kernels that produce their operands: warp shuffle ->
 * per-CTA shared reduce -> one atomicAdd(double) per CTA into a device scalar.
 */
#include "internal.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define SPMV_THREADS 128
/* Resident CTAs per SM the register allocator must leave room for.  Without
 * it ptxas squeezes the kernel into 32 registers (full-occupancy default) and
 * serialises the x gathers; shared memory limits residency to ~10 CTAs of 128
 * threads anyway, so 48 registers cost no occupancy. */
#define SPMV_MINB(T) ((T) >= 512 ? 2 : 1280 / (T))
#define SPMV_MAX_STAGES 8
#define BLAS1_THREADS 512

/* ------------------------------------------------------------------------ */
/* PTX wrappers: mbarrier + bulk async copy (TMA, 1-D)                       */
/* ------------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_init_fence()
{
    /* make the initialised barriers visible to the async (TMA) proxy */
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

/* global -> shared bulk copy; bytes and both addresses are multiples of 16.
 * The matrix streams are read exactly once per SpMV: evict-first keeps them
 * from displacing the x vector in L2. */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol) : "memory");
}

/* gather of a vector entry through the read-only path.  volatile: the batch of
 * gathers of one row stays a batch (the compiler may not sink individual loads
 * down to their uses, which would serialise their latencies). */
__device__ __forceinline__ double ld_x(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

/* coalesced read of matrix data that is used exactly once per SpMV: no L1 line, evict-first in L2,
 * so the stream does not displace the x vector (which every row re-reads) from either cache */
__device__ __forceinline__ double ld_stream(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

/* Programmatic dependent launch (opt-in, ACGB200_PDL=1).  Every kernel of the
 * iteration chain starts with this: wait until the preceding kernel of the
 * stream has completed and flushed (nothing in global memory is touched before),
 * then let the next kernel's CTAs be scheduled as soon as SM resources free up,
 * so its launch latency and CTA ramp overlap this kernel's tail.  Both
 * instructions are no-ops for a kernel launched without the attribute. */
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

/* ------------------------------------------------------------------------ */
/* reductions                                                                */
/* ------------------------------------------------------------------------ */

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int G>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* sum over the CTA, result valid in thread 0; `red` holds blockDim/32 doubles */
__device__ __forceinline__ double block_sum(double v, double *red)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}

/* ------------------------------------------------------------------------ */
/* iteration control                                                         */
/* ------------------------------------------------------------------------ */

struct Gate { int iter; bool active; };

/* Read the incoming control word.  The word a kernel reads is never written
 * by that same kernel, so every thread sees the same value. */
__device__ __forceinline__ Gate gate_read(const acgb200_ctrl *cin, const acgb200_devstate *st)
{
    Gate g; g.iter = 0; g.active = true;
    if (cin) {
        const int4 c = *reinterpret_cast<const int4 *>(cin);
        g.iter = c.x;
        g.active = (c.y == 0) && (c.x < st->maxits);
    }
    return g;
}

/* ------------------------------------------------------------------------ */
/* peer-memory exchange: flags and reductions                                */
/* ------------------------------------------------------------------------ */

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

#define RED_IDX(ch, parity, rank) ((((ch) * 2 + (parity)) * ACGB200_MAXR + (rank)) * 2)

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

/* Spin until *f >= seq.  A peer that never publishes (it failed, or a protocol
 * error) must not hang the GPU: after timeout_ns the waiter raises the sticky
 * timed_out flag of its descriptor and gives up, every later wait of the solve
 * returns at once, and the host turns the flag into an error (the iterates are
 * garbage from then on).  The clock is only read every 1024 polls. */
__device__ __forceinline__ void p2p_spin(const unsigned long long *f, unsigned long long seq, const acgb200_p2pdev *P)
{
    unsigned long long t0 = 0;
    unsigned int polls = 0;
    volatile unsigned long long *flag = const_cast<volatile unsigned long long *>(&P->timed_out);
    while (ld_acquire_sys(f) < seq) {
        if ((++polls & 1023u) != 0) continue;
        if (*flag) return;
        if (P->timeout_ns == 0) continue;
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > P->timeout_ns) { *flag = 1ull; __threadfence_system(); return; }
    }
}

/* Block-wide: wait until every sender has published halo sequence `seq`. */
__device__ __forceinline__ void p2p_wait_halo(const acgb200_p2pdev *P, unsigned long long seq)
{
    if ((int) threadIdx.x < P->nsenders) p2p_spin(P->my_hflag + P->senders[threadIdx.x], seq, P);
    __syncthreads();
}

/* Block-wide: wait until every rank has published sequence `seq` on channel
 * `ch`, then sum the partials of that parity in rank order (the same order on
 * every rank, so all ranks get bit-identical sums).  out[0..1] in shared memory. */
__device__ __forceinline__ void p2p_reduce(const acgb200_p2pdev *P, int ch, int parity, unsigned long long seq, double *out)
{
    if ((int) threadIdx.x < P->nranks) p2p_spin(P->my_rflag + ch * ACGB200_MAXR + threadIdx.x, seq, P);
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < P->nranks; r++) {
            const volatile double *s = P->my_red + RED_IDX(ch, parity, r);
            a += s[0]; b += s[1];
        }
        out[0] = a; out[1] = b;
    }
    __syncthreads();
}

/* sum of an already complete parity slot (no waiting) */
__device__ __forceinline__ double p2p_sum_slot(const acgb200_p2pdev *P, int ch, int parity)
{
    double a = 0.0;
    for (int r = 0; r < P->nranks; r++) a += ((const volatile double *) P->my_red)[RED_IDX(ch, parity, r)];
    return a;
}

/* True in exactly one CTA of the grid: the last one to get here.  Every CTA
 * must have fenced its own stores (system scope) before calling. */
__device__ __forceinline__ bool p2p_last_block(acgb200_p2pdev *P, int *flag_smem)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&P->ticket, 1u);
        *flag_smem = (t == gridDim.x - 1);
        if (*flag_smem) P->ticket = 0;
    }
    __syncthreads();
    const bool last = *flag_smem != 0;
    if (last) __threadfence_system();
    return last;
}

/* last CTA: write `count` reduction partials read from `src` (device scalars
 * completed by atomics of all CTAs) into every rank's slot, then the flags */
__device__ __forceinline__ void p2p_publish_red(acgb200_p2pdev *P, int ch, int parity, unsigned long long seq,
                                                double *src, int count)
{
    if ((int) threadIdx.x < P->nranks) {
        double *dst = P->peer_red[threadIdx.x] + RED_IDX(ch, parity, P->rank);
        dst[0] = atomicAdd(&src[0], 0.0);
        dst[1] = count > 1 ? atomicAdd(&src[1], 0.0) : 0.0;
        __threadfence_system();
        st_release_sys(P->peer_rflag[threadIdx.x] + ch * ACGB200_MAXR + P->rank, seq);
    }
}

__device__ __forceinline__ void p2p_publish_halo(acgb200_p2pdev *P, unsigned long long seq)
{
    if ((int) threadIdx.x < P->nrecip) st_release_sys(P->peer_hflag[threadIdx.x], seq);
}

/* a producer thread pushes the new value of border row `row` to the neighbours */
__device__ __forceinline__ void p2p_push_row(const acgb200_p2pdev *P, int row, int parity, double v)
{
    const int b = row - P->borderoff;
    for (int e = P->bptr[b]; e < P->bptr[b + 1]; e++) P->peer_ghost[P->bq[e]][parity][P->bdst[e]] = v;
}

/* Push halo values and/or reduction partials into the peers' windows, then
 * publish the sequence numbers (last block to finish does the publishing). */
__global__ void __launch_bounds__(512)
comm_post_kernel(const acgb200_postargs A)
{
    __shared__ int is_last;
    acgb200_p2pdev *P = A.p2p;
    int iter = A.iter_override;
    if (iter < 0) {
        const Gate g = gate_read(A.cin, A.st);
        if (!g.active) return;
        iter = g.iter;
    }
    if (A.vec) {
        const int parity = iter & 1;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P->sendsize; i += gridDim.x * blockDim.x) {
            int q = 0;
            while (i >= P->sdispls[q + 1]) q++;
            P->peer_ghost[q][parity][P->peer_rdispl[q] + (i - P->sdispls[q])] = A.vec[A.sendbufidx[i]];
        }
    }
    if (A.ch >= 0 && blockIdx.x == 0 && (int) threadIdx.x < P->nranks) {
        const int parity = (iter + A.par_off) & 1;
        const double *src = A.redbase + parity * A.redstride;
        double *dst = P->peer_red[threadIdx.x] + RED_IDX(A.ch, parity, P->rank);
        dst[0] = src[0];
        dst[1] = A.redcount > 1 ? src[1] : 0.0;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&P->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if (A.vec && (int) threadIdx.x < P->nrecip)
        st_release_sys(P->peer_hflag[threadIdx.x], P->hbase + (unsigned long long) iter);
    if (A.ch >= 0 && (int) threadIdx.x < P->nranks)
        st_release_sys(P->peer_rflag[threadIdx.x] + A.ch * ACGB200_MAXR + P->rank,
                       P->rbase + (unsigned long long) (iter + A.seq_off));
    if (threadIdx.x == 0) P->ticket = 0;
}

/* ------------------------------------------------------------------------ */
/* SpMV over TMA-staged row tiles                                            */
/* ------------------------------------------------------------------------ */

struct SpmvParams {
    const acgb200_tile *tiles;
    int ntiles;
    int sc;                 /* slots per stage for values/indices (multiple of 4) */
    int stage_bytes;
    int nstages;
    const int *rowptr;
    const int *colidx;
    const double *a;
    const double *x;
    double *y;
    const double *b;
    double *acc;
    int dotrows;
    int mode;
    const acgb200_ctrl *ctrl_in;
    acgb200_ctrl *ctrl_out;
    acgb200_devstate *st;
    int housekeeping;
    acgb200_p2pdev *p2p;
    int od_rowoffset, od_nrows;
    const int *orowptr;
    const int *ocolidx;
    const double *oa;
    int pub_ch;
};

__device__ __forceinline__ void spmv_issue(
    const SpmvParams &P, const acgb200_tile &tl, unsigned char *stage, uint64_t *bar, uint64_t pol)
{
    const int row_al = tl.row_begin & ~3;
    const int nrp = (tl.row_begin + tl.nrows + 1 - row_al + 3) & ~3;
    double *vals = reinterpret_cast<double *>(stage);
    int *cols = reinterpret_cast<int *>(stage + (size_t) P.sc * 8);
    int *rptr = reinterpret_cast<int *>(stage + (size_t) P.sc * 12);
    mbar_arrive_expect_tx(bar, (uint32_t) tl.nnz_al * 12u + (uint32_t) nrp * 4u);
    if (tl.nnz_al > 0) {
        bulk_g2s(vals, P.a + tl.k_al, (uint32_t) tl.nnz_al * 8u, bar, pol);
        bulk_g2s(cols, P.colidx + tl.k_al, (uint32_t) tl.nnz_al * 4u, bar, pol);
    }
    bulk_g2s(rptr, P.rowptr + row_al, (uint32_t) nrp * 4u, bar, pol);
}

/*
 * G lanes per row, T threads per CTA, U gathers in flight per lane.  A tile of
 * R rows is processed in ceil(R / (T/G)) passes; in a pass every G-lane group
 * owns one row and walks it in batches of U*G nonzeros: all U column indices
 * and values are read from shared memory first, then the U x-gathers are
 * issued back to back, then the FMAs -- the only long-latency operation (the
 * gather, an L1/L2 access) is thus U-deep per lane.  Slots past the end of the
 * row multiply a zero value with x[row's first column], which is always a
 * valid address.
 */
template <int G, int T, int U>
__global__ void __launch_bounds__(T, SPMV_MINB(T))
spmv_tiles_kernel(const SpmvParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[SPMV_MAX_STAGES];
    __shared__ double red[T / 32];
    __shared__ int last_flag;

    const int tid = threadIdx.x;
    pdl_prologue();
    const Gate gate = gate_read(P.ctrl_in, P.st);
    if (blockIdx.x == 0 && tid == 0 && P.ctrl_in) {
        *P.ctrl_out = *P.ctrl_in;
        if (gate.active) {
            /* zero the accumulator the *next* producer will add into; nothing
             * reads that slot while this kernel runs (control ring) */
            const int s = gate.iter & 1;
            if (P.housekeeping == 1) P.st->rr_loc[s ^ 1] = 0.0;
            if (P.housekeeping == 2) { P.st->gd_loc[s ^ 1][0] = 0.0; P.st->gd_loc[s ^ 1][1] = 0.0; }
        }
    }
    if (!gate.active) return;

    const int S = P.nstages;
    uint64_t pol = 0;
    if (tid == 0) {
        pol = l2_policy_evict_first();
        for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
        mbar_init_fence();
        for (int s = 0; s < S; s++) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < P.ntiles) spmv_issue(P, P.tiles[t], smem + (size_t) s * P.stage_bytes, &full_bar[s], pol);
        }
    }
    __syncthreads();

    constexpr int RPP = T / G;                 /* rows per pass */
    const int lane = tid % G;
    const int grp = tid / G;
    double dot = 0.0;
    const double *xg = NULL;            /* ghost values (peer-memory mode), set on the first border tile */

    int i = 0;
    for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x, i++) {
        const int s = i % S;
        const acgb200_tile tl = P.tiles[t];
        if (P.p2p && !xg && tl.row_begin + tl.nrows > P.od_rowoffset) {
            /* first tile with border rows: from here on the neighbours' values
             * are needed.  Border rows come last, so by now they have long
             * arrived and the wait is a single flag read per sender. */
            p2p_wait_halo(P.p2p, P.p2p->hbase + (unsigned long long) gate.iter);
            xg = P.p2p->my_ghost[gate.iter & 1] - P.od_nrows;
        }
        unsigned char *stage = smem + (size_t) s * P.stage_bytes;
        const double *vals = reinterpret_cast<const double *>(stage);
        const int *cols = reinterpret_cast<const int *>(stage + (size_t) P.sc * 8);
        const int *rp = reinterpret_cast<const int *>(stage + (size_t) P.sc * 12) + (tl.row_begin & 3);

        mbar_wait(&full_bar[s], (uint32_t) ((i / S) & 1));

        for (int base = 0; base < tl.nrows; base += RPP) {
            const int lr = base + grp;
            double sum = 0.0;
            if (lr < tl.nrows) {
                const int kb = rp[lr] - tl.k_al;
                const int ke = rp[lr + 1] - tl.k_al;
                for (int k = kb + lane; k < ke; k += U * G) {
                    int c[U];
                    double v[U], xv[U];
                    /* slots past the row end re-read its last entry and are
                     * zeroed after the load: straight-line code, no predicated
                     * loads, so all U gathers are issued before the first FMA */
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int kk = min(k + u * G, ke - 1);
                        c[u] = cols[kk];
                        v[u] = vals[kk];
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) xv[u] = ld_x(P.x + c[u]);
#pragma unroll
                    for (int u = 0; u < U; u++) sum = fma((k + u * G < ke) ? v[u] : 0.0, xv[u], sum);
                }
                if (xg && tl.row_begin + lr >= P.od_rowoffset) {
                    /* border x ghost block (acg/cgcuda.c:878), ocolidx rebased by
                     * -borderrowoffset so ghosts start at od_nrows */
                    const int ob = tl.row_begin + lr - P.od_rowoffset;
                    for (int k = P.orowptr[ob] + lane; k < P.orowptr[ob + 1]; k += G)
                        sum = fma(P.oa[k], xg[P.ocolidx[k]], sum);
                }
            }
            if (G > 1) sum = group_sum<G>(sum);
            if (lr < tl.nrows && lane == 0) {
                const int row = tl.row_begin + lr;
                if (P.mode == SPMV_R_B_AX) {
                    const double v = P.b[row] - sum;
                    P.y[row] = v;
                    if (row < P.dotrows) dot = fma(v, v, dot);
                } else {
                    P.y[row] = sum;
                    if (P.mode == SPMV_Y_AX_DOT && row < P.dotrows) dot = fma(__ldg(P.x + row), sum, dot);
                }
            }
        }

        __syncthreads();        /* every thread is done reading stage s */
        if (tid == 0) {
            const int tn = t + S * gridDim.x;
            if (tn < P.ntiles) spmv_issue(P, P.tiles[tn], stage, &full_bar[s], pol);
        }
    }

    if (P.acc) {
        const double v = block_sum(dot, red);
        if (tid == 0 && v != 0.0) atomicAdd(P.acc, v);
    }
    if (P.p2p && P.pub_ch >= 0 && P.p2p->fuse) {
        /* the last CTA to finish sends this rank's share of the fused dot to all ranks */
        __threadfence();
        if (p2p_last_block(P.p2p, &last_flag))
            p2p_publish_red(P.p2p, P.pub_ch, gate.iter & 1, P.p2p->rbase + (unsigned long long) gate.iter + 1ull, P.acc, 1);
    }
}

/* ------------------------------------------------------------------------ */
/* SpMV over pattern slices (slices.c): index-free, slice-major values         */
/* ------------------------------------------------------------------------ */

struct SliceParams {
    const acgb200_slice *slices;
    int nslices;
    const double *sval;
    const unsigned short *patid;   /* pattern id per row; ACGB200_NOPATTERN: exception row (EXC kernels only) */
    const int *spatoff;            /* [npat * lpad], zero beyond a pattern's length */
    int npat, lpad;
    const int *rowptr;             /* CSR row pointers / column indices: the columns of exception rows */
    const int *colidx;
    const double *x;
    double *y;
    const double *b;
    double *acc;
    int dotrows;
    int mode;
    const acgb200_ctrl *ctrl_in;
    acgb200_ctrl *ctrl_out;        /* NULL: the tile kernel launched behind this one forwards the word */
    acgb200_devstate *st;
    int housekeeping;
};

/*
 * One warp per slice, one row per lane.  Entry slot e of the slice's 32 rows is
 * one coalesced 256-byte load (ld_stream); the column of lane's entry is
 * row + offset[pattern of the row][e] from a zero-padded table in shared memory,
 * so padded slots (value 0) gather x[row].  UB value loads and UB gathers are
 * issued back to back before the UB FMAs; a row's products are added in CSR
 * order into one accumulator.  No shared-memory staging of the matrix: residency
 * is limited by registers only, the loads in flight per SM (warps x UB x 256 B
 * of values) cover the HBM latency.
 */
/* xr = x + row.  (Forming the column in 32-bit arithmetic first -- one IADD + IMAD.WIDE instead of the
 * sign extension and 64-bit add below -- was measured slower: 0.405 against 0.397 ms at C3, profiles/r02.) */
template <int UB>
__device__ __forceinline__ void slice_load(double (&vv)[UB], double (&xv)[UB], const double *v, const double *xr,
                                           const int *offs, int e, uint64_t pol)
{
#pragma unroll
    for (int u = 0; u < UB; u++) vv[u] = ld_stream(v + (size_t) (e + u) * 32, pol);
#pragma unroll
    for (int u = 0; u < UB; u++) xv[u] = ld_x(xr + offs[e + u]);
}

template <int UB>
__device__ __forceinline__ double slice_fma(const double (&vv)[UB], const double (&xv)[UB], double sum)
{
#pragma unroll
    for (int u = 0; u < UB; u++) sum = fma(vv[u], xv[u], sum);
    return sum;
}

/* Shapes measured and dropped in round 2 (profiles/r02/b_ab_224.log, c_ab_224.log; C3, this shape 0.397 ms):
 * prefetching the next batch into a second register set (0.450 ms: 64-96 registers cost more warps than the
 * prefetch buys), register caps of 40 / 32 per thread (spills), 32-bit column arithmetic (0.405), batches of
 * 14 (0.433), 256- and 64-thread CTAs (0.425 / 0.431), 3-5 wide batches (0.42-0.59).
 *
 * EXC: the plan has slices with exception rows (rows outside the dictionary, slices.c).  A warp whose 32 rows are
 * all in the dictionary runs the same loop as the plain kernel; otherwise the exception lanes take each slot's
 * column from the CSR index array (one extra load per slot for those lanes only), everything else is unchanged.
 * The plain instantiation is byte for byte the kernel measured in round 2. */
template <int UB, int T, bool EXC>
__global__ void __launch_bounds__(T, EXC ? 1280 / T : 0)
spmv_slices_kernel(const SliceParams P)
{
    extern __shared__ __align__(16) int spat_s[];
    __shared__ double red[T / 32];
    const int tid = threadIdx.x;
    pdl_prologue();
    const Gate gate = gate_read(P.ctrl_in, P.st);
    if (P.ctrl_out && blockIdx.x == 0 && tid == 0 && P.ctrl_in) {
        *P.ctrl_out = *P.ctrl_in;
        if (gate.active) {
            const int s = gate.iter & 1;
            if (P.housekeeping == 1) P.st->rr_loc[s ^ 1] = 0.0;
            if (P.housekeeping == 2) { P.st->gd_loc[s ^ 1][0] = 0.0; P.st->gd_loc[s ^ 1][1] = 0.0; }
        }
    }
    if (!gate.active) return;
    for (int i = tid; i < P.npat * P.lpad; i += T) spat_s[i] = P.spatoff[i];
    __syncthreads();

    const uint64_t pol = l2_policy_evict_first();
    const int lane = tid & 31;
    const int warp = (blockIdx.x * T + tid) >> 5;
    const int nwarps = (gridDim.x * T) >> 5;
    double dot = 0.0;
    for (int s = warp; s < P.nslices; s += nwarps) {
        const int4 sl = __ldg(reinterpret_cast<const int4 *>(P.slices) + s);     /* row0, nrows, len, vblk */
        const int row = sl.x + lane;
        const int pid = P.patid[row];
        const bool exc = EXC && pid == (int) ACGB200_NOPATTERN;
        const int *offs = spat_s + (exc ? 0 : pid) * P.lpad;
        const double *v = P.sval + ((size_t) sl.w << 5) + lane;
        const double *xr = P.x + row;
        const int L = sl.z;
        double sum = 0.0;
        int e = 0;
        if (EXC && __any_sync(0xffffffffu, exc)) {
            /* a slice with exception rows: those lanes take their columns from the index array (slots past the
             * row's end hold the value 0 and gather x[row]).  Branch-free: every lane issues the index load --
             * the lanes that do not need it read colidx[0], one cached sector for the whole warp -- and selects. */
            int kb = 0, len = 1;
            if (exc) { kb = P.rowptr[row]; len = P.rowptr[row + 1] - kb; }
            const int *cbase = P.colidx + kb;
            const int last = exc ? max(len - 1, 0) : 0;
            for (; e + UB <= L; e += UB) {
                double vv[UB], xv[UB];
                int cc[UB];
#pragma unroll
                for (int u = 0; u < UB; u++) vv[u] = ld_stream(v + (size_t) (e + u) * 32, pol);
#pragma unroll
                for (int u = 0; u < UB; u++) cc[u] = __ldg(cbase + min(e + u, last));
#pragma unroll
                for (int u = 0; u < UB; u++) {
                    const int col = exc ? (e + u < len ? cc[u] : row) : row + offs[e + u];
                    xv[u] = ld_x(P.x + col);
                }
                sum = slice_fma<UB>(vv, xv, sum);
            }
            for (; e < L; e++) {                                   /* the last, partial batch slot by slot */
                const int c1 = __ldg(cbase + min(e, last));
                const int col = exc ? (e < len ? c1 : row) : row + offs[e];
                sum = fma(ld_stream(v + (size_t) e * 32, pol), ld_x(P.x + col), sum);
            }
        }
        for (; e + UB <= L; e += UB) {
            double vv[UB], xv[UB];
            slice_load<UB>(vv, xv, v, xr, offs, e, pol);
            sum = slice_fma<UB>(vv, xv, sum);
        }
        if (e < L) {
            /* last, partial batch: slots past L are neither loaded nor added */
            double vv[UB], xv[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) vv[u] = e + u < L ? ld_stream(v + (size_t) (e + u) * 32, pol) : 0.0;
#pragma unroll
            for (int u = 0; u < UB; u++) xv[u] = e + u < L ? ld_x(xr + offs[e + u]) : 0.0;
#pragma unroll
            for (int u = 0; u < UB; u++) if (e + u < L) sum = fma(vv[u], xv[u], sum);
        }
        if (P.mode == SPMV_R_B_AX) {
            const double r = P.b[row] - sum;
            P.y[row] = r;
            if (row < P.dotrows) dot = fma(r, r, dot);
        } else {
            P.y[row] = sum;
            if (P.mode == SPMV_Y_AX_DOT && row < P.dotrows) dot = fma(__ldg(P.x + row), sum, dot);
        }
    }
    if (P.acc) {
        const double v = block_sum(dot, red);
        if (tid == 0 && v != 0.0) atomicAdd(P.acc, v);
    }
}

/* slice-major copy of the covered rows' values out of the device CSR arrays (once, at init) */
__global__ void __launch_bounds__(256)
slices_fill_kernel(int nslices, const acgb200_slice *slices, const int *__restrict__ rowptr,
                   const double *__restrict__ a, double *__restrict__ sval)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nslices; s += nwarps) {
        const acgb200_slice sl = slices[s];
        const int row = sl.row0 + lane;
        const int kb = rowptr[row], len = rowptr[row + 1] - kb;
        double *v = sval + ((size_t) sl.vblk << 5) + lane;
        for (int e = 0; e < sl.len; e++) v[(size_t) e * 32] = e < len ? a[kb + e] : 0.0;
    }
}

/* ------------------------------------------------------------------------ */
/* SpMV over merge-path tiles (mergeplan.c): power-law row lengths             */
/* ------------------------------------------------------------------------ */

struct MergeParams {
    const acgb200_mtile *tiles;
    int ntiles;
    int sc, rc;                    /* value/index slots and row-pointer slots per stage (multiples of 4) */
    int stage_bytes, nstages;
    const int *rowptr;
    const int *colidx;
    const double *a;
    double *part;                  /* [2 * ntiles] */
    const double *x;
    double *y;
    const double *b;
    double *acc;
    int dotrows;
    int mode;
    const acgb200_ctrl *ctrl_in;
    acgb200_ctrl *ctrl_out;        /* NULL: the tile kernel launched behind this one forwards the word */
    acgb200_devstate *st;
    int housekeeping;
};

#define MERGE_LONGSEG 32           /* pieces longer than this are summed by a whole warp */
#define MERGE_LISTCAP 128

__device__ __forceinline__ void merge_issue(const MergeParams &P, const acgb200_mtile &tl, unsigned char *stage,
                                            uint64_t *bar, uint64_t pol)
{
    const int k_al = tl.k0 & ~3;
    const int nnz_al = (tl.k0 + tl.nnz - k_al + 3) & ~3;
    const int row_al = tl.r0 & ~3;
    const int nrp = (tl.r0 + tl.nre + 1 - row_al + 3) & ~3;
    double *vals = reinterpret_cast<double *>(stage);
    int *cols = reinterpret_cast<int *>(stage + (size_t) P.sc * 8);
    int *rptr = reinterpret_cast<int *>(stage + (size_t) P.sc * 12);
    mbar_arrive_expect_tx(bar, (uint32_t) nnz_al * 12u + (uint32_t) nrp * 4u);
    if (nnz_al > 0) {
        bulk_g2s(vals, P.a + k_al, (uint32_t) nnz_al * 8u, bar, pol);
        bulk_g2s(cols, P.colidx + k_al, (uint32_t) nnz_al * 4u, bar, pol);
    }
    bulk_g2s(rptr, P.rowptr + row_al, (uint32_t) nrp * 4u, bar, pol);
}

/* what a finished row does with its sum (same epilogues as the tile kernel) */
__device__ __forceinline__ void spmv_row_epilogue(int row, double sum, int mode, const double *x, double *y, const double *b,
                                                  int dotrows, double &dot)
{
    if (mode == SPMV_R_B_AX) {
        const double v = b[row] - sum;
        y[row] = v;
        if (row < dotrows) dot = fma(v, v, dot);
    } else {
        y[row] = sum;
        if (mode == SPMV_Y_AX_DOT && row < dotrows) dot = fma(__ldg(x + row), sum, dot);
    }
}

/*
 * Per tile: (1) every thread multiplies a strided share of the tile's nonzeros -- U gathers in
 * flight, equal work for all threads whatever the rows look like -- and leaves the products in
 * shared memory; (2) the tile's units (one per row end, plus the piece behind the last row end) are
 * summed from shared memory: a thread per unit, units longer than MERGE_LONGSEG by a warp.  A unit
 * that is a whole row gets the epilogue; the piece of a split row goes to part[] for the fix-up
 * kernel (slot 0: the row ends here but began in an earlier tile; slot 1: the row continues).
 */
template <int T, int U>
__global__ void __launch_bounds__(T)
spmv_merge_kernel(const MergeParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[SPMV_MAX_STAGES];
    __shared__ double red[T / 32];
    __shared__ int lst[MERGE_LISTCAP];
    __shared__ int nlst;

    const int tid = threadIdx.x;
    pdl_prologue();
    const Gate gate = gate_read(P.ctrl_in, P.st);
    if (P.ctrl_out && blockIdx.x == 0 && tid == 0 && P.ctrl_in) {
        *P.ctrl_out = *P.ctrl_in;
        if (gate.active) {
            const int s = gate.iter & 1;
            if (P.housekeeping == 1) P.st->rr_loc[s ^ 1] = 0.0;
            if (P.housekeeping == 2) { P.st->gd_loc[s ^ 1][0] = 0.0; P.st->gd_loc[s ^ 1][1] = 0.0; }
        }
    }
    if (!gate.active) return;

    const int S = P.nstages;
    double *prod = reinterpret_cast<double *>(smem + (size_t) S * P.stage_bytes);     /* [sc] products of the current tile */
    uint64_t pol = 0;
    if (tid == 0) {
        nlst = 0;
        pol = l2_policy_evict_first();
        for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
        mbar_init_fence();
        for (int s = 0; s < S; s++) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < P.ntiles) merge_issue(P, P.tiles[t], smem + (size_t) s * P.stage_bytes, &full_bar[s], pol);
        }
    }
    __syncthreads();

    double dot = 0.0;
    int i = 0;
    for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x, i++) {
        const int s = i % S;
        const acgb200_mtile tl = P.tiles[t];
        unsigned char *stage = smem + (size_t) s * P.stage_bytes;
        const int koff = tl.k0 & 3;                    /* first own nonzero inside the 16-byte aligned slice */
        const double *vals = reinterpret_cast<const double *>(stage) + koff;
        const int *cols = reinterpret_cast<const int *>(stage + (size_t) P.sc * 8) + koff;
        const int *rp = reinterpret_cast<const int *>(stage + (size_t) P.sc * 12) + (tl.r0 & 3);    /* rp[j] = rowptr[r0 + j] */

        mbar_wait(&full_bar[s], (uint32_t) ((i / S) & 1));

        /* (1) products */
        for (int base = 0; base < tl.nnz; base += T * U) {
            int c[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int kk = min(base + u * T + tid, tl.nnz - 1);
                c[u] = cols[kk];
                v[u] = vals[kk];
            }
#pragma unroll
            for (int u = 0; u < U; u++) xv[u] = ld_x(P.x + c[u]);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int kk = base + u * T + tid;
                if (kk < tl.nnz) prod[kk] = v[u] * xv[u];
            }
        }
        __syncthreads();

        /* (2a) a thread per unit: j < nre is row r0 + j, j == nre the piece behind the last row end */
        const int kend = tl.k0 + tl.nnz;
        for (int j = tid; j <= tl.nre; j += T) {
            const int a0 = max(rp[j], tl.k0) - tl.k0;
            const int b0 = (j < tl.nre ? rp[j + 1] : kend) - tl.k0;
            if (b0 - a0 > MERGE_LONGSEG) {
                const int slot = atomicAdd(&nlst, 1);
                if (slot < MERGE_LISTCAP) { lst[slot] = j; continue; }
                /* (cannot happen: a tile holds fewer than MERGE_LISTCAP pieces of that length; summed here if it did) */
            }
            double sum = 0.0;
            for (int k = a0; k < b0; k++) sum += prod[k];
            if (j == tl.nre) P.part[2 * (size_t) t + 1] = sum;
            else if (j == 0 && rp[0] < tl.k0) P.part[2 * (size_t) t] = sum;
            else spmv_row_epilogue(tl.r0 + j, sum, P.mode, P.x, P.y, P.b, P.dotrows, dot);
        }
        if (tid == 0 && !(tl.nre > 0 && rp[0] < tl.k0)) P.part[2 * (size_t) t] = 0.0;      /* no head piece in this tile */
        __syncthreads();

        /* (2b) long units: a warp each, lanes stride over the piece, fixed shuffle tree */
        const int nl = min(nlst, MERGE_LISTCAP);
        for (int q = tid >> 5; q < nl; q += T / 32) {
            const int j = lst[q];
            const int a0 = max(rp[j], tl.k0) - tl.k0;
            const int b0 = (j < tl.nre ? rp[j + 1] : kend) - tl.k0;
            double sum = 0.0;
            for (int k = a0 + (tid & 31); k < b0; k += 32) sum += prod[k];
            sum = warp_sum(sum);
            if ((tid & 31) == 0) {
                if (j == tl.nre) P.part[2 * (size_t) t + 1] = sum;
                else if (j == 0 && rp[0] < tl.k0) P.part[2 * (size_t) t] = sum;
                else spmv_row_epilogue(tl.r0 + j, sum, P.mode, P.x, P.y, P.b, P.dotrows, dot);
            }
        }
        __syncthreads();        /* stage s, prod[] and the list are free again */
        if (tid == 0) {
            nlst = 0;
            const int tn = t + S * gridDim.x;
            if (tn < P.ntiles) merge_issue(P, P.tiles[tn], stage, &full_bar[s], pol);
        }
    }

    if (P.acc) {
        const double v = block_sum(dot, red);
        if (tid == 0 && v != 0.0) atomicAdd(P.acc, v);
    }
}

/* split rows: add the pieces in tile order (a warp per row), then the row's epilogue */
__global__ void __launch_bounds__(256)
spmv_merge_fix_kernel(int nsplit, const acgb200_msplit *split, const double *part,
                      const double *x, double *y, const double *b, double *acc, int dotrows, int mode,
                      const acgb200_ctrl *cin, const acgb200_devstate *st)
{
    __shared__ double red[8];
    if (!gate_read(cin, st).active) return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (int i = warp; i < nsplit; i += nwarps) {
        const acgb200_msplit sp = split[i];
        double sum = 0.0;
        for (int t = sp.ta + lane; t < sp.tb; t += 32) sum += part[2 * (size_t) t + 1];
        sum = warp_sum(sum);
        if (lane == 0) {
            sum += part[2 * (size_t) sp.tb];
            spmv_row_epilogue(sp.row, sum, mode, x, y, b, dotrows, dot);
        }
    }
    if (acc) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0 && dot != 0.0) atomicAdd(acc, dot);
    }
}

/* ---- long rows: several CTAs per row, partials to scratch, then a finisher -- */

__global__ void __launch_bounds__(256)
spmv_long_partial_kernel(int nlong, const int *longrows, int chunks,
                         const int *rowptr, const int *colidx, const double *a, const double *x,
                         double *scratch, const acgb200_ctrl *cin, const acgb200_devstate *st)
{
    __shared__ double red[8];
    if (!gate_read(cin, st).active) return;
    const int lrow = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int row = longrows[lrow];
    const long long kb = rowptr[row], ke = rowptr[row + 1];
    const long long len = ke - kb;
    const long long c0 = kb + len * chunk / chunks, c1 = kb + len * (chunk + 1) / chunks;
    double sum = 0.0;
    for (long long k = c0 + threadIdx.x; k < c1; k += blockDim.x)
        sum = fma(a[k], __ldg(x + colidx[k]), sum);
    sum = block_sum(sum, red);
    if (threadIdx.x == 0) scratch[blockIdx.x] = sum;
}

__global__ void spmv_long_finish_kernel(int nlong, const int *longrows, int chunks, const double *scratch,
                                        const double *x, double *y, const double *b, double *acc,
                                        int dotrows, int mode, const acgb200_ctrl *cin, const acgb200_devstate *st,
                                        const acgb200_p2pdev *p2p, int od_rowoffset, int od_nrows,
                                        const int *orowptr, const int *ocolidx, const double *oa)
{
    const Gate g = gate_read(cin, st);
    if (!g.active) return;
    /* peer-memory mode: long rows that are border rows also get their
     * border x ghost entries here (the tile kernel only sees rows in tiles) */
    const double *xg = NULL;
    if (p2p) {
        p2p_wait_halo(p2p, p2p->hbase + (unsigned long long) g.iter);
        xg = p2p->my_ghost[g.iter & 1] - od_nrows;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double dot = 0.0;
    if (i < nlong) {
        const int row = longrows[i];
        double sum = 0.0;
        for (int c = 0; c < chunks; c++) sum += scratch[(size_t) i * chunks + c];
        if (xg && row >= od_rowoffset) {
            const int ob = row - od_rowoffset;
            for (int k = orowptr[ob]; k < orowptr[ob + 1]; k++) sum = fma(oa[k], xg[ocolidx[k]], sum);
        }
        if (mode == SPMV_R_B_AX) {
            const double v = b[row] - sum; y[row] = v;
            if (row < dotrows) dot = v * v;
        } else {
            y[row] = sum;
            if (mode == SPMV_Y_AX_DOT && row < dotrows) dot = x[row] * sum;
        }
    }
    dot = warp_sum(dot);
    if (acc && (threadIdx.x & 31) == 0 && dot != 0.0) atomicAdd(acc, dot);
}

/* ---- medium rows: one warp per row, straight from global memory (opt-in) ----- */

/*
 * Rows that would monopolise one G-lane group of a tile (a few hundred to a
 * couple of thousand nonzeros, the body of a power-law degree distribution) are
 * taken out of the tiles by the planner when option "spmv_medium" is set and
 * processed here: a warp walks its row with coalesced 32-wide accesses, four
 * gathers in flight per lane, and applies the same epilogue as the tile kernel.
 * Persistent grid, warps stride over the list; one atomic per CTA for the dot.
 */
__global__ void __launch_bounds__(256)
spmv_medium_kernel(int nmed, const int *medrows,
                   const int *rowptr, const int *colidx, const double *a, const double *x,
                   double *y, const double *b, double *acc, int dotrows, int mode,
                   const acgb200_ctrl *cin, const acgb200_devstate *st,
                   const acgb200_p2pdev *p2p, int od_rowoffset, int od_nrows,
                   const int *orowptr, const int *ocolidx, const double *oa)
{
    __shared__ double red[8];
    const Gate g = gate_read(cin, st);
    if (!g.active) return;
    const double *xg = NULL;
    if (p2p) {
        p2p_wait_halo(p2p, p2p->hbase + (unsigned long long) g.iter);
        xg = p2p->my_ghost[g.iter & 1] - od_nrows;
    }
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (int i = warp; i < nmed; i += nwarps) {
        const int row = medrows[i];
        const int kb = rowptr[row], ke = rowptr[row + 1];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = kb + lane;
        for (; k + 96 < ke; k += 128) {
            const double x0 = __ldg(x + colidx[k]), x1 = __ldg(x + colidx[k + 32]);
            const double x2 = __ldg(x + colidx[k + 64]), x3 = __ldg(x + colidx[k + 96]);
            s0 = fma(a[k], x0, s0); s1 = fma(a[k + 32], x1, s1);
            s2 = fma(a[k + 64], x2, s2); s3 = fma(a[k + 96], x3, s3);
        }
        for (; k < ke; k += 32) s0 = fma(a[k], __ldg(x + colidx[k]), s0);
        double sum = (s0 + s1) + (s2 + s3);
        if (xg && row >= od_rowoffset) {
            const int ob = row - od_rowoffset;
            for (int j = orowptr[ob] + lane; j < orowptr[ob + 1]; j += 32) sum = fma(oa[j], xg[ocolidx[j]], sum);
        }
        sum = warp_sum(sum);
        if (lane == 0) {
            if (mode == SPMV_R_B_AX) {
                const double v = b[row] - sum;
                y[row] = v;
                if (row < dotrows) dot = fma(v, v, dot);
            } else {
                y[row] = sum;
                if (mode == SPMV_Y_AX_DOT && row < dotrows) dot = fma(x[row], sum, dot);
            }
        }
    }
    if (acc) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0 && dot != 0.0) atomicAdd(acc, dot);
    }
}

/* ---- border x ghost block: thread per border row, direct from global -------- */

__global__ void __launch_bounds__(256)
offdiag_kernel(const acgb200_offdiagargs A)
{
    __shared__ double red[8];
    const Gate g = gate_read(A.ctrl_in, A.st);
    if (!g.active) return;
    /* ghost values: the tail of x (filled by ncclRecv) or, with the peer-memory
     * exchange, this iteration's parity buffer of the window, valid once every
     * sender has published the iteration's sequence number.  ocolidx is rebased
     * by -borderrowoffset (acg/symcsrmatrix.c:838), ghosts start at nrows. */
    const double *xg = A.x + A.rowoffset;
    if (A.p2p) {
        const int it = A.p2p_iter_override >= 0 ? A.p2p_iter_override : g.iter;
        p2p_wait_halo(A.p2p, A.p2p->hbase + (unsigned long long) it);
        xg = A.p2p->my_ghost[it & 1] - A.nrows;
    }
    double dot = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.nrows; i += gridDim.x * blockDim.x) {
        const int kb = A.orowptr[i], ke = A.orowptr[i + 1];
        const int row = A.rowoffset + i;
        double sum = 0.0;
        for (int k = kb; k < ke; k++) sum = fma(A.oa[k], xg[A.ocolidx[k]], sum);
        double v = A.y[row];
        if (kb != ke) { v = A.minus ? v - sum : v + sum; A.y[row] = v; }
        if (A.dotkind == 1) dot = fma(__ldg(A.x + row), v, dot);
        else if (A.dotkind == 2) dot = fma(v, v, dot);
    }
    if (A.acc) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0 && dot != 0.0) atomicAdd(A.acc, dot);
    }
}

/* ------------------------------------------------------------------------ */
/* fused BLAS-1                                                              */
/* ------------------------------------------------------------------------ */

/* r -= alpha t with alpha = (r,r)/(p,Ap), and the new (r,r) in the same pass */
__global__ void __launch_bounds__(BLAS1_THREADS)
cg_update_r_kernel(int n, acgb200_devstate *st, int cin, int cout, int multi, acgb200_p2pdev *P,
                   const double *__restrict__ t, double *__restrict__ r)
{
    __shared__ double red[BLAS1_THREADS / 32];
    __shared__ double glob[2];
    __shared__ int last_flag;
    pdl_prologue();
    const Gate g = gate_read(&st->ctrl[cin], st);
    if (cin != cout && blockIdx.x == 0 && threadIdx.x == 0) st->ctrl[cout] = st->ctrl[cin];
    if (!g.active) return;
    const int s = g.iter & 1;
    double rr, pap;
    if (P) {
        /* (p,Ap) of this iteration arrives on channel 0; (r,r) was completed on
         * channel 1 one iteration ago (setup value for the first iteration) */
        p2p_reduce(P, 0, s, P->rbase + (unsigned long long) g.iter + 1ull, glob);
        pap = glob[0];
        rr = g.iter > 0 ? p2p_sum_slot(P, 1, s) : st->rr[0];
    } else {
        rr = multi ? st->rr[s] : st->rr_loc[s];
        pap = multi ? st->pap[s] : st->pap_loc[s];
    }
    const double alpha = rr / pap;
    double acc = 0.0;
    const int n2 = n >> 1;
    const double2 *t2 = reinterpret_cast<const double2 *>(t);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 tv = t2[i];
        double2 rv = r2[i];
        rv.x = fma(-alpha, tv.x, rv.x);
        rv.y = fma(-alpha, tv.y, rv.y);
        r2[i] = rv;
        acc = fma(rv.x, rv.x, acc);
        acc = fma(rv.y, rv.y, acc);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double rv = fma(-alpha, t[n - 1], r[n - 1]);
        r[n - 1] = rv;
        acc = fma(rv, rv, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(&st->rr_loc[s ^ 1], acc);
    if (P && P->fuse) {
        __threadfence();
        if (p2p_last_block(P, &last_flag))
            p2p_publish_red(P, 1, s ^ 1, P->rbase + (unsigned long long) g.iter + 1ull, &st->rr_loc[s ^ 1], 1);
    }
}

/* x += alpha p ; p = r + beta p ; decides convergence for the next iteration */
__global__ void __launch_bounds__(BLAS1_THREADS)
cg_update_xp_kernel(int n, acgb200_devstate *st, int cin, int cout, int multi, acgb200_p2pdev *P,
                    const double *__restrict__ r, double *__restrict__ p, double *__restrict__ x)
{
    __shared__ double glob[2];
    __shared__ int last_flag;
    pdl_prologue();
    const Gate g = gate_read(&st->ctrl[cin], st);
    const int s = g.iter & 1;
    double rr, rrn, pap;
    if (P && g.active) {
        p2p_reduce(P, 1, s ^ 1, P->rbase + (unsigned long long) g.iter + 1ull, glob);
        rrn = glob[0];
        rr = g.iter > 0 ? p2p_sum_slot(P, 1, s) : st->rr[0];
        pap = p2p_sum_slot(P, 0, s);
    } else {
        rr = multi ? st->rr[s] : st->rr_loc[s];
        rrn = multi ? st->rr[s ^ 1] : st->rr_loc[s ^ 1];
        pap = multi ? st->pap[s] : st->pap_loc[s];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        acgb200_ctrl c = st->ctrl[cin];
        if (g.active) {
            c.iter = g.iter + 1;
            if (P) st->rr[s ^ 1] = rrn;    /* keep the global value where the host reads it */
            /* acg/cgcuda.c:1008-1012: test ||r|| < tol with the norm, strictly */
            if (st->tol > 0.0 && sqrt(rrn) < st->tol) { c.done = 1; st->final_rr = rrn; }
            st->pap_loc[s ^ 1] = 0.0;      /* accumulator of the next SpMV */
        }
        st->ctrl[cout] = c;
    }
    if (!g.active) return;
    const double alpha = rr / pap;
    const double beta = rrn / rr;
    const int n2 = n >> 1;
    const double2 *r2 = reinterpret_cast<const double2 *>(r);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    double2 *x2 = reinterpret_cast<double2 *>(x);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const double2 rv = r2[i];
        double2 pv = p2[i];
        double2 xv = x2[i];
        xv.x = fma(alpha, pv.x, xv.x);
        xv.y = fma(alpha, pv.y, xv.y);
        pv.x = fma(beta, pv.x, rv.x);
        pv.y = fma(beta, pv.y, rv.y);
        x2[i] = xv;
        p2[i] = pv;
        if (P && P->fuse && 2 * i + 1 >= P->borderoff) {
            /* p is the input of the next SpMV: its border entries go straight
             * into the neighbours' ghost buffers of the next iteration's parity */
            if (2 * i >= P->borderoff) p2p_push_row(P, 2 * i, s ^ 1, pv.x);
            p2p_push_row(P, 2 * i + 1, s ^ 1, pv.y);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double pv = p[n - 1];
        x[n - 1] = fma(alpha, pv, x[n - 1]);
        const double pn = fma(beta, pv, r[n - 1]);
        p[n - 1] = pn;
        if (P && P->fuse && n - 1 >= P->borderoff) p2p_push_row(P, n - 1, s ^ 1, pn);
    }
    if (P && P->fuse) {
        __threadfence_system();
        if (p2p_last_block(P, &last_flag)) p2p_publish_halo(P, P->hbase + (unsigned long long) g.iter + 1ull);
    }
}

/* Pipelined CG: z=q+beta z; t=w+beta t; p=r+beta p; x+=alpha p; r-=alpha t;
 * w-=alpha z (acg/cg-kernels-cuda.cu:201-214), plus gamma'=(r,r), delta'=(w,r)
 * of the updated vectors for the next iteration. */
/* (Measured and dropped in round 2, profiles/r02/d_ab_112.log, e3_ab_224_n2.log: two rows per thread and trip
 * -- 8 % faster at one rank's share of the 8-GPU problem, 5 % slower at C3, nothing between two GPUs; the
 * system fence moved behind the border rows -- no effect.) */
__global__ void __launch_bounds__(BLAS1_THREADS)
pcg_update_kernel(int n, acgb200_devstate *st, int cin, int cout, int multi, acgb200_p2pdev *P,
                  const double *__restrict__ q, double *__restrict__ z, double *__restrict__ w,
                  double *__restrict__ t, double *__restrict__ p, double *__restrict__ r,
                  double *__restrict__ x)
{
    __shared__ double red[BLAS1_THREADS / 32];
    __shared__ double glob[2];
    __shared__ int last_flag;
    pdl_prologue();
    const Gate g = gate_read(&st->ctrl[cin], st);
    const int s = g.iter & 1;
    double gamma, delta;
    if (P && g.active && g.iter > 0) {
        /* {gamma,delta} pushed by every rank after its previous update */
        p2p_reduce(P, 0, s, P->rbase + (unsigned long long) g.iter, glob);
        gamma = glob[0]; delta = glob[1];
    } else {
        gamma = multi ? st->gd[s][0] : st->gd_loc[s][0];
        delta = multi ? st->gd[s][1] : st->gd_loc[s][1];
    }
    const double gamma_prev = st->prev[s][0], alpha_prev = st->prev[s][1];
    /* acg/cgcuda.c:1764-1772: the test precedes the update */
    const bool conv = st->tol > 0.0 && sqrt(gamma) < st->tol;
    const double beta = gamma / gamma_prev;
    const double alpha = gamma / (delta - beta * gamma / alpha_prev);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        acgb200_ctrl c = st->ctrl[cin];
        if (g.active) {
            if (P) { st->gd[s][0] = gamma; st->gd[s][1] = delta; }   /* for the host's final report */
            if (conv) { c.done = 1; st->final_rr = gamma; }
            else { c.iter = g.iter + 1; st->prev[s ^ 1][0] = gamma; st->prev[s ^ 1][1] = alpha; }
        }
        st->ctrl[cout] = c;
    }
    if (!g.active || conv) return;
    double g2 = 0.0, d2 = 0.0;
    /* Peer-memory mode: border rows first.  Their new w values are stored into
     * the neighbours' ghost buffers (w is the input of the next SpMV), and
     * issuing those NVLink stores at the start lets them drain under the
     * interior rows instead of in front of the closing system fence. */
    const bool push = P && P->fuse;
    const int first = push ? P->borderoff : 0;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
    for (int pass = 0; pass < 2; pass++) {
        const int lo = pass == 0 ? first : 0, hi = pass == 0 ? n : first;
        for (int i = lo + gtid; i < hi; i += gstride) {
            const double zv = fma(beta, z[i], q[i]);
            const double tv = fma(beta, t[i], w[i]);
            const double pv = fma(beta, p[i], r[i]);
            const double rv = fma(-alpha, tv, r[i]);
            const double wv = fma(-alpha, zv, w[i]);
            z[i] = zv; t[i] = tv; p[i] = pv;
            x[i] = fma(alpha, pv, x[i]);
            r[i] = rv;
            w[i] = wv;
            g2 = fma(rv, rv, g2);
            d2 = fma(wv, rv, d2);
            if (push && pass == 0) p2p_push_row(P, i, s ^ 1, wv);
        }
    }
    g2 = block_sum(g2, red);
    d2 = block_sum(d2, red);
    if (threadIdx.x == 0) {
        atomicAdd(&st->gd_loc[s ^ 1][0], g2);
        atomicAdd(&st->gd_loc[s ^ 1][1], d2);
    }
    if (P && P->fuse) {
        /* last CTA: this rank's {gamma,delta} of the next iteration to every
         * rank, and the halo sequence number to the neighbours */
        __threadfence_system();
        if (p2p_last_block(P, &last_flag)) {
            const unsigned long long it1 = (unsigned long long) g.iter + 1ull;
            p2p_publish_red(P, 0, s ^ 1, P->rbase + it1, &st->gd_loc[s ^ 1][0], 2);
            p2p_publish_halo(P, P->hbase + it1);
        }
    }
}

__global__ void __launch_bounds__(BLAS1_THREADS)
dot_kernel(int n, const double *__restrict__ x, const double *__restrict__ y, double *acc)
{
    __shared__ double red[BLAS1_THREADS / 32];
    double v = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v = fma(x[i], y[i], v);
    v = block_sum(v, red);
    if (threadIdx.x == 0) atomicAdd(acc, v);
}

__global__ void __launch_bounds__(BLAS1_THREADS)
dot2_kernel(int n, const double *__restrict__ r, const double *__restrict__ w, double *acc2)
{
    __shared__ double red[BLAS1_THREADS / 32];
    double g = 0.0, d = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double rv = r[i];
        g = fma(rv, rv, g);
        d = fma(w[i], rv, d);
    }
    g = block_sum(g, red);
    d = block_sum(d, red);
    if (threadIdx.x == 0) { atomicAdd(&acc2[0], g); atomicAdd(&acc2[1], d); }
}

__global__ void gather_kernel(int n, double *__restrict__ dst, const double *__restrict__ src, const int *__restrict__ idx)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void scatter_kernel(int n, const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ idx)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[idx[i]] = src[i];
}

/* ------------------------------------------------------------------------ */
/* host-side launchers (C linkage)                                           */
/* ------------------------------------------------------------------------ */

static int g_num_sms = 0;

extern "C" int acgb200_num_sms(void)
{
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0)
            g_num_sms = 148;
    }
    return g_num_sms;
}

/* CTAs per SM for the BLAS-1 kernels: 0 = as many as are resident at once */
static int g_blas1_ctas_per_sm = 0;

extern "C" void acgb200_blas1_set_ctas_per_sm(int v) { g_blas1_ctas_per_sm = v < 0 ? 0 : v; }


/* One wave of grid-stride CTAs: SMs x resident CTAs of that kernel (a grid that
 * is not a multiple of it ends in a partial wave at a fraction of the memory
 * parallelism), fewer when the vector is short.  `slot` caches the occupancy. */
static int blas1_grid(int n, const void *fn, int *slot)
{
    if (*slot == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, BLAS1_THREADS, 0) != cudaSuccess || per_sm < 1)
            per_sm = 2;
        *slot = per_sm;
    }
    const int per = BLAS1_THREADS * 2;
    long long want = ((long long) n + per - 1) / per;
    const long long cap = (long long) acgb200_num_sms() * (g_blas1_ctas_per_sm > 0 ? g_blas1_ctas_per_sm : *slot);
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int) want;
}

/* kernel launch, with the programmatic-dependency attribute when PDL is on */
static int g_pdl = 0;

extern "C" void acgb200_set_pdl(int v) { g_pdl = v != 0; }

template <typename... Params, typename... Args>
static cudaError_t launch_chain(void (*kernel)(Params...), int grid, int block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned) grid); lc.blockDim = dim3((unsigned) block);
    lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kernel, args...);
}

typedef void (*spmv_fn)(const SpmvParams);

/* instantiated (lanes per row, threads per CTA) pairs; gathers in flight U=8 */
#define SPMV_VARIANTS(X) \
    X(1, 64) X(2, 64) X(4, 64) X(8, 128) X(16, 128) X(32, 128) X(1, 128) X(1, 256) X(2, 128) X(2, 256) X(2, 512) X(4, 128) X(4, 256) X(4, 512) \
    X(8, 256) X(8, 512) X(16, 256) X(16, 512) X(32, 256) X(32, 512)

static spmv_fn spmv_variant(int G, int T, int U)
{
    if (U == 4 && G == 1 && T == 128) return spmv_tiles_kernel<1, 128, 4>;
    if (U == 16 && G == 1 && T == 128) return spmv_tiles_kernel<1, 128, 16>;
    if (U == 16 && G == 1 && T == 256) return spmv_tiles_kernel<1, 256, 16>;
    if (U == 4 && G == 2 && T == 256) return spmv_tiles_kernel<2, 256, 4>;
    if (U == 16 && G == 2 && T == 256) return spmv_tiles_kernel<2, 256, 16>;
    if (U == 4 && G == 4 && T == 512) return spmv_tiles_kernel<4, 512, 4>;
    if (U == 4 && G == 4 && T == 128) return spmv_tiles_kernel<4, 128, 4>;
    if (U == 4 && G == 4 && T == 256) return spmv_tiles_kernel<4, 256, 4>;
#define X(g, t) if (G == g && T == t) return spmv_tiles_kernel<g, t, 8>;
    SPMV_VARIANTS(X)
#undef X
    return NULL;
}

static inline int stage_slots(const acgb200_spmvplan *pl) { return (pl->nnz_cap + 8 + 3) & ~3; }
static inline int stage_rslots(const acgb200_spmvplan *pl) { return (pl->rows_cap + 1 + 8 + 3) & ~3; }
static inline int stage_bytes(const acgb200_spmvplan *pl)
{
    const int sc = stage_slots(pl);
    const int rc = stage_rslots(pl);
    return (sc * 12 + rc * 4 + 127) & ~127;
}

typedef void (*slice_fn)(const SliceParams);

/* (values + gathers in flight per lane, threads per CTA) of the slice kernel */
static slice_fn slice_variant(int UB, int T, int EXC)
{
#define X(u, t) if (UB == u && T == t) return EXC ? spmv_slices_kernel<u, t, true> : spmv_slices_kernel<u, t, false>;
    X(5, 128) X(7, 128) X(8, 128) X(9, 128) X(7, 256) X(8, 256) X(9, 256)
#undef X
    return NULL;
}

typedef void (*merge_fn)(const MergeParams);

static merge_fn merge_variant(int T)
{
    if (T == 128) return spmv_merge_kernel<128, 8>;
    if (T == 256) return spmv_merge_kernel<256, 4>;
    return NULL;
}

static inline int merge_slots(const acgb200_spmvplan *pl) { return (pl->merge_items + 8 + 3) & ~3; }
static inline int merge_rslots(const acgb200_spmvplan *pl) { return (pl->merge_items + 1 + 8 + 3) & ~3; }
static inline int merge_stage_bytes(const acgb200_spmvplan *pl) { return (merge_slots(pl) * 12 + merge_rslots(pl) * 4 + 127) & ~127; }

/* acgb200_spmv_choose (tile-plan heuristic) and acgb200_spmv_min_bytes are host-only: plan.c */

extern "C" int acgb200_spmv_configure(acgb200_spmvplan *pl)
{
    const void *fn = (const void *) spmv_variant(pl->lanes_per_row, pl->threads, pl->unroll);
    if (!fn) return (int) cudaErrorInvalidConfiguration;
    pl->smem_bytes = stage_bytes(pl) * pl->nstages;
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem_bytes);
    if (err) return (int) err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, pl->threads, pl->smem_bytes);
    if (err) return (int) err;
    if (per_sm < 1) per_sm = 1;
    if (pl->max_ctas_per_sm > 0 && per_sm > pl->max_ctas_per_sm) per_sm = pl->max_ctas_per_sm;
    long long grid = (long long) acgb200_num_sms() * per_sm;
    if (grid > pl->ntiles) grid = pl->ntiles;
    if (grid < 1) grid = 1;
    pl->grid = (int) grid;
    if (pl->nmtiles > 0) {
        merge_fn mfn = merge_variant(pl->merge_threads);
        if (!mfn) return (int) cudaErrorInvalidConfiguration;
        pl->merge_smem = merge_stage_bytes(pl) * pl->merge_stages + merge_slots(pl) * (int) sizeof(double);
        err = cudaFuncSetAttribute((const void *) mfn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->merge_smem);
        if (err) return (int) err;
        int mper = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mper, (const void *) mfn, pl->merge_threads, pl->merge_smem);
        if (err) return (int) err;
        if (mper < 1) mper = 1;
        if (pl->merge_max_ctas > 0 && mper > pl->merge_max_ctas) mper = pl->merge_max_ctas;
        long long mgrid = (long long) acgb200_num_sms() * mper;
        if (mgrid > pl->nmtiles) mgrid = pl->nmtiles;
        pl->merge_grid = (int) (mgrid < 1 ? 1 : mgrid);
    }
    if (pl->nslices > 0) {
        slice_fn sfn = slice_variant(pl->slice_ub, pl->slice_threads, pl->slice_exc > 0);
        if (!sfn) return (int) cudaErrorInvalidConfiguration;
        pl->slice_smem = ((pl->slice_npat * pl->slice_lpad + 3) & ~3) * (int) sizeof(int);
        err = cudaFuncSetAttribute((const void *) sfn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->slice_smem);
        if (err) return (int) err;
        int sper = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sper, (const void *) sfn, pl->slice_threads, pl->slice_smem);
        if (err) return (int) err;
        if (sper < 1) sper = 1;
        if (pl->slice_max_ctas > 0 && sper > pl->slice_max_ctas) sper = pl->slice_max_ctas;
        long long sgrid = (long long) acgb200_num_sms() * sper;
        const long long need = ((long long) pl->nslices * 32 + pl->slice_threads - 1) / pl->slice_threads;
        if (sgrid > need) sgrid = need;
        pl->slice_grid = (int) (sgrid < 1 ? 1 : sgrid);
    }
    return 0;
}

extern "C" int acgb200_slices_fill(const acgb200_spmvplan *pl, const int *d_rowptr, const double *d_a, cudaStream_t stream)
{
    if (pl->nslices <= 0) return 0;
    long long grid = ((long long) pl->nslices * 32 + 255) / 256;
    const long long cap = (long long) acgb200_num_sms() * 8;
    if (grid > cap) grid = cap;
    slices_fill_kernel<<<(int) grid, 256, 0, stream>>>(pl->nslices, pl->d_slices, d_rowptr, d_a, pl->d_sval);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_spmv_launch(const acgb200_spmvargs *a, cudaStream_t stream)
{
    const acgb200_spmvplan *pl = a->plan;
    /* with neither tiles nor a peer-memory duty for the tile kernel, the slice kernel forwards the
     * control word itself and the tile kernel is not launched */
    const bool slices_forward = (pl->nslices > 0 || pl->nmtiles > 0) && pl->ntiles == 0 && !a->p2p;
    if (pl->nmtiles > 0) {
        /* rows [0, merge_rows) of an irregular matrix: merge-path tiles, then the rows cut by tile boundaries */
        MergeParams M;
        M.tiles = pl->d_mtiles; M.ntiles = pl->nmtiles; M.sc = merge_slots(pl); M.rc = merge_rslots(pl);
        M.stage_bytes = merge_stage_bytes(pl); M.nstages = pl->merge_stages;
        M.rowptr = a->rowptr; M.colidx = a->colidx; M.a = a->a; M.part = pl->d_mpart;
        M.x = a->x; M.y = a->y; M.b = a->b; M.acc = a->acc; M.dotrows = a->dotrows; M.mode = a->mode;
        M.ctrl_in = a->ctrl_in; M.ctrl_out = slices_forward ? a->ctrl_out : NULL; M.st = a->st; M.housekeeping = a->housekeeping;
        const cudaError_t le = launch_chain(merge_variant(pl->merge_threads), pl->merge_grid, pl->merge_threads,
                                            (size_t) pl->merge_smem, stream, M);
        if (le) return (int) le;
        if (pl->nsplit > 0) {
            long long grid = ((long long) pl->nsplit * 32 + 255) / 256;
            const long long cap = (long long) acgb200_num_sms() * 8;
            if (grid > cap) grid = cap;
            /* reads the control word the merge kernel read (ctrl_in is only rewritten by the iteration's last kernel) */
            spmv_merge_fix_kernel<<<(int) grid, 256, 0, stream>>>(pl->nsplit, pl->d_msplit, pl->d_mpart, a->x, a->y, a->b, a->acc,
                                                                 a->dotrows, a->mode, a->ctrl_in, a->st);
            const cudaError_t fe = cudaGetLastError();
            if (fe) return (int) fe;
        }
    }
    if (pl->nslices > 0) {
        /* first: the tile kernel's last CTA publishes the fused dot, which these rows add into */
        SliceParams S;
        S.slices = pl->d_slices; S.nslices = pl->nslices; S.sval = pl->d_sval;
        S.patid = pl->d_spatid; S.spatoff = pl->d_spatoff; S.npat = pl->slice_npat; S.lpad = pl->slice_lpad;
        S.rowptr = a->rowptr; S.colidx = a->colidx;
        S.x = a->x; S.y = a->y; S.b = a->b; S.acc = a->acc; S.dotrows = a->dotrows; S.mode = a->mode;
        S.ctrl_in = a->ctrl_in; S.ctrl_out = slices_forward ? a->ctrl_out : NULL; S.st = a->st; S.housekeeping = a->housekeeping;
        const cudaError_t le = launch_chain(slice_variant(pl->slice_ub, pl->slice_threads, pl->slice_exc > 0), pl->slice_grid, pl->slice_threads,
                                            (size_t) pl->slice_smem, stream, S);
        if (le) return (int) le;
    }
    if (pl->ntiles > 0 || (a->ctrl_in && !slices_forward)) {   /* with no tiles the kernel still forwards the control word */
        SpmvParams P;
        P.tiles = pl->d_tiles; P.ntiles = pl->ntiles;
        P.sc = stage_slots(pl); P.stage_bytes = stage_bytes(pl); P.nstages = pl->nstages;
        P.rowptr = a->rowptr; P.colidx = a->colidx; P.a = a->a; P.x = a->x; P.y = a->y; P.b = a->b;
        P.acc = a->acc; P.dotrows = a->dotrows; P.mode = a->mode;
        P.ctrl_in = a->ctrl_in; P.ctrl_out = a->ctrl_out; P.st = a->st; P.housekeeping = a->housekeeping;
        P.p2p = (acgb200_p2pdev *) a->p2p; P.od_rowoffset = a->od_rowoffset; P.od_nrows = a->od_nrows;
        P.orowptr = a->orowptr; P.ocolidx = a->ocolidx; P.oa = a->oa; P.pub_ch = a->p2p ? a->pub_ch : -1;
        const cudaError_t le = launch_chain(spmv_variant(pl->lanes_per_row, pl->threads, pl->unroll),
                                            pl->grid, pl->threads, (size_t) pl->smem_bytes, stream, P);
        if (le) return (int) le;
        cudaError_t err = cudaGetLastError();
        if (err) return (int) err;
    }
    if (pl->nmed > 0) {
        long long grid = ((long long) pl->nmed + 7) / 8;
        const long long cap = (long long) acgb200_num_sms() * 8;
        if (grid > cap) grid = cap;
        spmv_medium_kernel<<<(int) grid, 256, 0, stream>>>(
            pl->nmed, pl->d_medrows, a->rowptr, a->colidx, a->a, a->x, a->y, a->b, a->acc,
            a->dotrows, a->mode, a->ctrl_in, a->st,
            a->p2p, a->od_rowoffset, a->od_nrows, a->orowptr, a->ocolidx, a->oa);
        cudaError_t err = cudaGetLastError();
        if (err) return (int) err;
    }
    if (pl->nlong > 0) {
        if (!pl->d_long_scratch) return (int) cudaErrorInvalidValue;
        spmv_long_partial_kernel<<<pl->nlong * pl->long_chunks, 256, 0, stream>>>(
            pl->nlong, pl->d_longrows, pl->long_chunks, a->rowptr, a->colidx, a->a, a->x, pl->d_long_scratch,
            a->ctrl_in, a->st);
        spmv_long_finish_kernel<<<(pl->nlong + 127) / 128, 128, 0, stream>>>(
            pl->nlong, pl->d_longrows, pl->long_chunks, pl->d_long_scratch, a->x, a->y, a->b, a->acc,
            a->dotrows, a->mode, a->ctrl_in, a->st,
            a->p2p, a->od_rowoffset, a->od_nrows, a->orowptr, a->ocolidx, a->oa);
        cudaError_t err = cudaGetLastError();
        if (err) return (int) err;
    }
    return 0;
}

extern "C" int acgb200_offdiag_launch(const acgb200_offdiagargs *a, cudaStream_t stream)
{
    if (a->nrows <= 0) return 0;
    int grid = (a->nrows + 255) / 256;
    const int cap = acgb200_num_sms() * 8;
    if (grid > cap) grid = cap;
    offdiag_kernel<<<grid, 256, 0, stream>>>(*a);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_cg_update_r(int n, acgb200_devstate *st, int cin, int cout, int multi,
                                   acgb200_p2pdev *p2p,
                                   const double *t, double *r, cudaStream_t stream)
{
    static int occ = 0;
    return (int) launch_chain(cg_update_r_kernel, blas1_grid(n, (const void *) cg_update_r_kernel, &occ), BLAS1_THREADS, 0, stream,
                              n, st, cin, cout, multi, p2p, t, r);
}

extern "C" int acgb200_cg_update_xp(int n, acgb200_devstate *st, int cin, int cout, int multi,
                                    acgb200_p2pdev *p2p,
                                    const double *r, double *p, double *x, cudaStream_t stream)
{
    static int occ = 0;
    return (int) launch_chain(cg_update_xp_kernel, blas1_grid(n, (const void *) cg_update_xp_kernel, &occ), BLAS1_THREADS, 0, stream,
                              n, st, cin, cout, multi, p2p, r, p, x);
}

extern "C" int acgb200_pcg_update(int n, acgb200_devstate *st, int cin, int cout, int multi,
                                  acgb200_p2pdev *p2p,
                                  const double *q, double *z, double *w, double *t, double *p,
                                  double *r, double *x, cudaStream_t stream)
{
    static int occ = 0;
    return (int) launch_chain(pcg_update_kernel, blas1_grid(n, (const void *) pcg_update_kernel, &occ), BLAS1_THREADS, 0, stream,
                              n, st, cin, cout, multi, p2p, q, z, w, t, p, r, x);
}

extern "C" int acgb200_comm_post(const acgb200_postargs *a, cudaStream_t stream)
{
    /* a handful of CTAs moves a halo of 10^4-10^5 doubles in a few microseconds;
     * a reduction-only post needs one */
    const int grid = a->vec ? 8 : 1;
    comm_post_kernel<<<grid, 512, 0, stream>>>(*a);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_dot(int n, const double *x, const double *y, double *acc, cudaStream_t stream)
{
    static int occ = 0;
    dot_kernel<<<blas1_grid(n, (const void *) dot_kernel, &occ), BLAS1_THREADS, 0, stream>>>(n, x, y, acc);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_dot2(int n, const double *r, const double *w, double *acc2, cudaStream_t stream)
{
    static int occ = 0;
    dot2_kernel<<<blas1_grid(n, (const void *) dot2_kernel, &occ), BLAS1_THREADS, 0, stream>>>(n, r, w, acc2);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_gather(int n, double *dst, const double *src, const int *idx, cudaStream_t stream)
{
    if (n <= 0) return 0;
    gather_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, dst, src, idx);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_scatter(int n, const double *src, double *dst, const int *idx, cudaStream_t stream)
{
    if (n <= 0) return 0;
    scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, src, dst, idx);
    return (int) cudaGetLastError();
}

/* ------------------------------------------------------------------------ */
/* BLAS-1 building blocks of acg/cg-kernels-cuda.h:45-97 (public, unused by   */
/* this library's own loops)                                                  */
/* ------------------------------------------------------------------------ */

enum { HELP_AXPY_QUOT = 0, HELP_AXMY_QUOT = 1, HELP_AYPX_QUOT = 2 };

/* y = (num/den) x + y | y = -(num/den) x + y | y = (num/den) y + x, the quotient read on the device */
template <int OP>
__global__ void __launch_bounds__(BLAS1_THREADS)
helper_axpy_kernel(int n, const double *num, const double *den, const double *__restrict__ x, double *__restrict__ y)
{
    const double a = OP == HELP_AXMY_QUOT ? -(*num) / (*den) : (*num) / (*den);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[i] = OP == HELP_AYPX_QUOT ? fma(a, y[i], x[i]) : fma(a, x[i], y[i]);
}

__global__ void helper_scalars_kernel(double *out0, double *out1, const double *num, const double *den, int op)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double q = (*num) / (*den);
    if (op == 0) { *out0 = q; *out1 = -q; }        /* alpha, -alpha */
    else *out0 = q;                                /* beta */
}

/* the recurrences of acg/cg-kernels-cuda.cu:201-214; the scalars gamma_prev, alpha_prev are updated by
 * the single-thread kernel launched behind it (stream order replaces the reference's grid sync) */
__global__ void __launch_bounds__(BLAS1_THREADS)
helper_pipelined_kernel(int n, const double *gamma, const double *gamma_prev, const double *delta, const double *alpha_prev,
                        const double *__restrict__ q, double *__restrict__ p, double *__restrict__ r, double *__restrict__ t,
                        double *__restrict__ x, double *__restrict__ z, double *__restrict__ w)
{
    const double beta = (*gamma) / (*gamma_prev);
    const double alpha = (*gamma) / ((*delta) - beta * (*gamma) / (*alpha_prev));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double zv = fma(beta, z[i], q[i]);
        const double tv = fma(beta, t[i], w[i]);
        const double pv = fma(beta, p[i], r[i]);
        z[i] = zv; t[i] = tv; p[i] = pv;
        x[i] = fma(alpha, pv, x[i]);
        r[i] = fma(-alpha, tv, r[i]);
        w[i] = fma(-alpha, zv, w[i]);
    }
}

__global__ void helper_pipelined_scalars_kernel(const double *gamma, double *gamma_prev, const double *delta, double *alpha_prev)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double beta = (*gamma) / (*gamma_prev);
    const double alpha = (*gamma) / ((*delta) - beta * (*gamma) / (*alpha_prev));
    *gamma_prev = *gamma;
    *alpha_prev = alpha;
}

extern "C" int acgb200_helper_axpy(int op, int n, const double *num, const double *den, const double *x, double *y, cudaStream_t stream)
{
    static int occ = 0;
    const int grid = blas1_grid(n, (const void *) helper_axpy_kernel<HELP_AXPY_QUOT>, &occ);
    if (op == HELP_AXPY_QUOT) helper_axpy_kernel<HELP_AXPY_QUOT><<<grid, BLAS1_THREADS, 0, stream>>>(n, num, den, x, y);
    else if (op == HELP_AXMY_QUOT) helper_axpy_kernel<HELP_AXMY_QUOT><<<grid, BLAS1_THREADS, 0, stream>>>(n, num, den, x, y);
    else helper_axpy_kernel<HELP_AYPX_QUOT><<<grid, BLAS1_THREADS, 0, stream>>>(n, num, den, x, y);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_helper_scalars(int op, double *out0, double *out1, const double *num, const double *den, cudaStream_t stream)
{
    helper_scalars_kernel<<<1, 32, 0, stream>>>(out0, out1, num, den, op);
    return (int) cudaGetLastError();
}

extern "C" int acgb200_helper_pipelined(int n, const double *gamma, double *gamma_prev, const double *delta, const double *q,
                                        double *p, double *r, double *t, double *x, double *z, double *w, double *alpha_prev,
                                        cudaStream_t stream)
{
    static int occ = 0;
    const int grid = blas1_grid(n, (const void *) helper_pipelined_kernel, &occ);
    helper_pipelined_kernel<<<grid, BLAS1_THREADS, 0, stream>>>(n, gamma, gamma_prev, delta, alpha_prev, q, p, r, t, x, z, w);
    helper_pipelined_scalars_kernel<<<1, 32, 0, stream>>>(gamma, gamma_prev, delta, alpha_prev);
    return (int) cudaGetLastError();
}
