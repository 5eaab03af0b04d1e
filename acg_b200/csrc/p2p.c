/*
 * p2p.c -- set-up of the peer-memory exchange (CUDA IPC over NVLink/NVSwitch).
 *
 * One process per GPU; the only bootstrap channel is the NCCL communicator the
 * caller passed in: the IPC handle of this rank's window and the layout facts
 * its peers need are all-gathered with ncclAllGather.  See internal.h for the
 * window layout and kernels.cu (comm_post_kernel, p2p_wait_halo, p2p_reduce)
 * for the data path.
 */
#include "acgb200/error.h"
#include "acgb200/halo.h"
#include "internal.h"
#include "p2p.h"

#include <stdlib.h>
#include <string.h>

struct record {                       /* what every rank tells every other rank */
    cudaIpcMemHandle_t handle;
    int valid;                        /* 0: this rank could not set up its window -- nobody uses the exchange */
    int recvsize;
    int rdispl_for[ACGB200_MAXR];     /* offset of sender q's segment in my ghost buffer, -1: not a neighbour */
};

static size_t off_hflag(void) { return 0; }
static size_t off_rflag(void) { return off_hflag() + ACGB200_MAXR * sizeof(unsigned long long); }
static size_t off_red(void) { return off_rflag() + ACGB200_NCH * ACGB200_MAXR * sizeof(unsigned long long); }
static size_t off_ghost(void) { return off_red() + (size_t) ACGB200_NCH * 2 * ACGB200_MAXR * 2 * sizeof(double); }
static size_t ghost_stride(int recvsize) { return (((size_t) (recvsize > 0 ? recvsize : 1) + 15) & ~(size_t) 15) * sizeof(double); }

/*
 * Inverse of the halo send list: for border row b (relative to borderoff) the
 * entries e in [bptr[b], bptr[b+1]) name the recipient index bq[e] and the
 * position bdst[e] in that recipient's ghost buffer (its rdispl for this sender
 * plus the offset inside the segment).  Host-only, no CUDA.
 */
int acgb200_p2p_inverse_map(const struct acghalo *halo, int borderoff, int nborder,
                            const int *rdispl_at_recipient, int *bptr, int *bq, int *bdst)
{
    int *cnt = calloc((size_t) nborder + 2, sizeof(int));
    if (!cnt) return ACG_ERR_ERRNO;
    for (int i = 0; i < halo->sendsize; i++) {
        const int b = halo->sendbufidx[i] - borderoff;
        if (b < 0 || b >= nborder) { free(cnt); return ACG_ERR_INVALID_VALUE; }
        cnt[b + 2]++;
    }
    for (int b = 0; b < nborder; b++) cnt[b + 2] += cnt[b + 1];
    for (int q = 0; q < halo->nrecipients; q++) {
        for (int i = halo->sdispls[q]; i < halo->sdispls[q] + halo->sendcounts[q]; i++) {
            const int b = halo->sendbufidx[i] - borderoff;
            const int e = cnt[b + 1]++;
            bq[e] = q;
            bdst[e] = rdispl_at_recipient[q] + (i - halo->sdispls[q]);
        }
    }
    memcpy(bptr, cnt, ((size_t) nborder + 1) * sizeof(int));
    free(cnt);
    return ACG_SUCCESS;
}

void acgb200_p2p_free(struct acgb200_p2p *p)
{
    if (!p) return;
    for (int r = 0; r < p->nranks; r++)
        if (r != p->rank && p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
    cudaFree((void *) p->h_desc.bptr); cudaFree((void *) p->h_desc.bq); cudaFree((void *) p->h_desc.bdst);
    cudaFree(p->d_desc);
    /* the window itself is freed by the caller after a barrier (peers may still
     * have it mapped) */
}

int acgb200_p2p_init(struct acgb200_p2p *p, const struct acghalo *halo, int borderoff, int nborder,
                     const struct acgcomm *comm, cudaStream_t stream, int *errcode)
{
    memset(p, 0, sizeof(*p));
    int nranks = 1, rank = 0;
    int err = acgcomm_size(comm, &nranks); if (err) return err;
    err = acgcomm_rank(comm, &rank); if (err) return err;
    if (nranks < 2 || nranks > ACGB200_MAXR || comm->type != acgcomm_nccl) return ACG_ERR_NOT_SUPPORTED;
    /* (every rank sees the same communicator, so the returns above are taken by all ranks or by none) */
    p->nranks = nranks; p->rank = rank;

    /* One error path: everything allocated here is released at `done` unless the set-up succeeds.  The
     * all-gather is executed by every rank, also by one whose local part failed (its record says so): a
     * rank that skipped the collective would leave its peers inside it and mismatch the communicator's
     * next collective. */
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (errcode) *errcode = (int) e_; err = ACG_ERR_CUDA; goto done; } } while (0)
    struct record mine, *all = NULL;
    void *d_mine = NULL, *d_all = NULL;
    int *bptr = NULL, *bq = NULL, *bdst = NULL, *d_bptr = NULL, *d_bq = NULL, *d_bdst = NULL;
    struct acgb200_p2pdev *d = &p->h_desc;
    memset(&mine, 0, sizeof(mine));
    mine.valid = 0;
    if (halo->nrecipients <= ACGB200_MAXR && halo->nsenders <= ACGB200_MAXR) {
        p->window_bytes = off_ghost() + 2 * ghost_stride(halo->recvsize);
        cudaError_t e = cudaMalloc(&p->window, p->window_bytes);
        if (!e) e = cudaMemset(p->window, 0, p->window_bytes);
        if (!e) e = cudaIpcGetMemHandle(&mine.handle, p->window);
        if (!e) mine.valid = 1;
        else { if (errcode) *errcode = (int) e; (void) cudaGetLastError(); }
    }
    mine.recvsize = halo->recvsize;
    for (int q = 0; q < ACGB200_MAXR; q++) mine.rdispl_for[q] = -1;
    if (mine.valid) for (int j = 0; j < halo->nsenders; j++) mine.rdispl_for[halo->senders[j]] = halo->rdispls[j];
    all = malloc((size_t) nranks * sizeof(*all));
    if (!all) { err = ACG_ERR_ERRNO; goto done; }      /* (host malloc of a few KiB: not a case worth a collective) */
    CUP(cudaMalloc(&d_mine, sizeof(mine)));
    CUP(cudaMalloc(&d_all, (size_t) nranks * sizeof(mine)));
    CUP(cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream));
    {
        ncclResult_t nr = ncclAllGather(d_mine, d_all, sizeof(mine), ncclChar, comm->ncclcomm, stream);
        if (nr != ncclSuccess) { if (errcode) *errcode = (int) nr; err = ACG_ERR_NCCL; goto done; }
    }
    CUP(cudaMemcpyAsync(all, d_all, (size_t) nranks * sizeof(mine), cudaMemcpyDeviceToHost, stream));
    CUP(cudaStreamSynchronize(stream));
    for (int r = 0; r < nranks; r++)
        if (!all[r].valid) { err = mine.valid ? ACG_ERR_NOT_SUPPORTED : ACG_ERR_CUDA; goto done; }   /* same verdict on every rank */

    /* map every peer's window */
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { p->peer_base[r] = p->window; continue; }
        CUP(cudaIpcOpenMemHandle(&p->peer_base[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess));
    }

    /* device descriptor */
    memset(d, 0, sizeof(*d));
    d->nranks = nranks; d->rank = rank;
    d->nrecip = halo->nrecipients; d->sendsize = halo->sendsize;
    for (int i = 0; i < halo->nrecipients; i++) {
        const int q = halo->recipients[i];
        d->sdispls[i] = halo->sdispls[i];
        d->peer_rdispl[i] = all[q].rdispl_for[rank];
        if (d->peer_rdispl[i] < 0) { err = ACG_ERR_INVALID_VALUE; goto done; }   /* asymmetric pattern */
        char *base = p->peer_base[q];
        d->peer_ghost[i][0] = (double *) (base + off_ghost());
        d->peer_ghost[i][1] = (double *) (base + off_ghost() + ghost_stride(all[q].recvsize));
        d->peer_hflag[i] = (unsigned long long *) (base + off_hflag()) + rank;
    }
    d->sdispls[halo->nrecipients] = halo->sendsize;
    d->nsenders = halo->nsenders;
    for (int j = 0; j < halo->nsenders; j++) d->senders[j] = halo->senders[j];
    {
        char *me = p->window;
        d->my_hflag = (unsigned long long *) (me + off_hflag());
        d->my_ghost[0] = (double *) (me + off_ghost());
        d->my_ghost[1] = (double *) (me + off_ghost() + ghost_stride(halo->recvsize));
        for (int r = 0; r < nranks; r++) {
            char *base = p->peer_base[r];
            d->peer_red[r] = (double *) (base + off_red());
            d->peer_rflag[r] = (unsigned long long *) (base + off_rflag());
        }
        d->my_red = (double *) (me + off_red());
        d->my_rflag = (unsigned long long *) (me + off_rflag());
    }
    d->hbase = d->rbase = 1;
    {
        /* watchdog of the device-side waits (kernels.cu, p2p_spin): generous, its
         * job is to turn a dead peer into an error instead of a hung GPU */
        const char *t = getenv("ACGB200_P2P_TIMEOUT_MS");
        const long long ms = t ? atoll(t) : 30000;
        d->timeout_ns = ms > 0 ? (unsigned long long) ms * 1000000ull : 0ull;
        d->timed_out = 0;
    }
    /* inverse send map: which (neighbour, ghost offset) pairs each border row feeds */
    d->borderoff = borderoff; d->nborder = nborder;
    {
        const size_t ne = (size_t) (halo->sendsize > 0 ? halo->sendsize : 1);
        bptr = malloc(((size_t) nborder + 1) * sizeof(int));
        bq = malloc(ne * sizeof(int));
        bdst = malloc(ne * sizeof(int));
        if (!bptr || !bq || !bdst) { err = ACG_ERR_ERRNO; goto done; }
        err = acgb200_p2p_inverse_map(halo, borderoff, nborder, d->peer_rdispl, bptr, bq, bdst);
        if (err) goto done;
        CUP(cudaMalloc((void **) &d_bptr, ((size_t) nborder + 1) * sizeof(int)));
        CUP(cudaMalloc((void **) &d_bq, ne * sizeof(int)));
        CUP(cudaMalloc((void **) &d_bdst, ne * sizeof(int)));
        CUP(cudaMemcpy(d_bptr, bptr, ((size_t) nborder + 1) * sizeof(int), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(d_bq, bq, (size_t) halo->sendsize * sizeof(int), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(d_bdst, bdst, (size_t) halo->sendsize * sizeof(int), cudaMemcpyHostToDevice));
    }
    CUP(cudaMalloc((void **) &p->d_desc, sizeof(*d)));
    d->bptr = d_bptr; d->bq = d_bq; d->bdst = d_bdst;
    CUP(cudaMemcpy(p->d_desc, d, sizeof(*d), cudaMemcpyHostToDevice));
    d_bptr = d_bq = d_bdst = NULL;             /* owned by the descriptor from here on (acgb200_p2p_free) */
    p->seq = 1;
    p->enabled = 1;
    err = ACG_SUCCESS;
done:
    free(all); free(bptr); free(bq); free(bdst);
    cudaFree(d_mine); cudaFree(d_all);
    cudaFree(d_bptr); cudaFree(d_bq); cudaFree(d_bdst);
    if (err) {
        /* the caller's agreement step (min-allreduce over the ranks' outcomes, cgcuda.c) follows on all ranks;
         * the window stays allocated until acgsolvercuda_free, after a barrier: a peer may have mapped it */
        for (int r = 0; r < nranks; r++)
            if (r != rank && p->peer_base[r]) { cudaIpcCloseMemHandle(p->peer_base[r]); p->peer_base[r] = NULL; }
        cudaFree(p->d_desc); p->d_desc = NULL;
        memset(&p->h_desc, 0, sizeof(p->h_desc));
        p->enabled = 0;
    }
    return err;
#undef CUP
}

/* start a new solve: fresh sequence bases above everything published so far
 * (identical on all ranks, which make identical call sequences) */
int acgb200_p2p_begin(struct acgb200_p2p *p, int maxits, cudaStream_t stream)
{
    p->seq += 4;
    p->h_desc.hbase = p->h_desc.rbase = p->seq;
    p->seq += (unsigned long long) maxits + 4;
    /* hbase, rbase and the watchdog flag are adjacent: one 24-byte copy */
    p->h_desc.timed_out = 0;
    cudaError_t e = cudaMemcpyAsync(&p->d_desc->hbase, &p->h_desc.hbase, 3 * sizeof(unsigned long long),
                                    cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    return e == cudaSuccess ? ACG_SUCCESS : ACG_ERR_CUDA;
}

/* after a solve: did any wait for a peer time out? */
int acgb200_p2p_timed_out(struct acgb200_p2p *p, cudaStream_t stream, int *flag)
{
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, &p->d_desc->timed_out, sizeof(v), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    *flag = v != 0;
    return e == cudaSuccess ? ACG_SUCCESS : ACG_ERR_CUDA;
}
