/*
 * acgb200/halo.h -- halo (ghost) exchange pattern and its device-side state.
 *
 * ABI counterpart of acg/halo.h:72-186 (struct acghalo) and :308-352 (struct
 * acghaloexchange as laid out with ACG_HAVE_CUDA and without ACG_HAVE_HIP),
 * and of the CUDA+NCCL entry points :370-527.  acgsolvercuda_fwritempi reads
 * cg->halo->{nexchanges,npack,...} and cg->haloexchange->maxevents
 * (acg/cgcuda.c:1976-1980, :2068-2069), hence the exact layouts.
 *
 * Pattern facts the device path relies on (acg/graph.c:1898-1981): the send
 * set is a gather from the border range [borderrowoffset, ghostrowoffset) of a
 * vector, one contiguous sendbuf segment per neighbour; the receive set is the
 * ghost tail.  When recvbufidx[k] == recvbufidx[0] + k (always the case for
 * patterns built from acggraph neighbours) the exchange receives straight into
 * the ghost tail and no unpack kernel runs.
 */
#ifndef ACGB200_HALO_H
#define ACGB200_HALO_H

#include "acgb200/config.h"
#include "acgb200/comm.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* acg/halo.h:72-186 */
struct acghalo {
    int nrecipients;
    int *recipients, *sendcounts, *sdispls;
    int sendsize;
    int *sendbufidx;
    int nsenders;
    int *senders, *recvcounts, *rdispls;
    int recvsize;
    int *recvbufidx;
    int nexchanges;
    double texchange;
    double tpack, tunpack, tsendrecv, tmpiirecv, tmpisend, tmpiwaitall;
    int64_t npack, nunpack, nmpiirecv, nmpisend;
    int64_t Bpack, Bunpack, Bmpiirecv, Bmpisend;
    int maxexchangestats;
    double (*thaloexchangestats)[4];
};

/* acg/halo.h:308-352 (CUDA variant) */
struct acghaloexchange {
    enum acgdatatype sendtype, recvtype;
    void *sendbuf, *recvbuf;
    void *sendreqs, *recvreqs;
    void *d_sendbuf, *d_recvbuf;
    void *d_sendbufidx, *d_recvbufidx;
    int *d_recipients, *d_sendcounts, *d_sdispls;
    int *d_senders, *d_recvcounts, *d_rdispls;
    int *d_putdispls, *d_putranks, *d_getranks;
    cudaStream_t cudastream;
    uint64_t *d_received, *d_readytoreceive;
    int *putdispls, *putranks, *getranks;
    int use_nvshmem, use_rocshmem;
    int maxevents, nevents;
    cudaEvent_t (*cudaevents)[4];
};

/* acg/halo.h:205 */
ACG_API void acghalo_free(struct acghalo *halo);

/* acg/halo.h:370 */
ACG_API int acghaloexchange_init_cuda(
    struct acghaloexchange *haloexchange, const struct acghalo *halo,
    enum acgdatatype sendtype, enum acgdatatype recvtype,
    const struct acgcomm *comm, cudaStream_t stream);
/* acg/halo.h:397 */
ACG_API void acghaloexchange_free(struct acghaloexchange *haloexchange);
/* acg/halo.h:404 */
ACG_API int acghaloexchange_profile(
    const struct acghaloexchange *haloexchange, int maxevents, int *nevents,
    double *texchange, double *tpack, double *tsendrecv, double *tunpack);

/* acg/halo.h:426, :449 -- gather / scatter kernels (acg/halo.cu:41, :94) */
ACG_API int acghalo_pack_cuda(
    int sendbufsize, void *d_sendbuf, enum acgdatatype datatype,
    int srcbufsize, const void *d_srcbuf, const int *d_srcbufidx,
    cudaStream_t stream, int64_t *nbytes, int *errcode);
ACG_API int acghalo_unpack_cuda(
    int recvbufsize, const void *d_recvbuf, enum acgdatatype datatype,
    int dstbufsize, void *d_dstbuf, const int *d_dstbufidx,
    cudaStream_t stream, int64_t *nbytes, int *errcode);

/* acg/halo.h:490, :513 -- begin: pack + grouped ncclSend/ncclRecv on `stream`
 * (acg/halo.c:1456, :1272); end: unpack (skipped when received in place)
 * (acg/halo.c:1552) */
ACG_API int acghalo_exchange_cuda_begin(
    struct acghalo *halo, struct acghaloexchange *haloexchange,
    int srcbufsize, const void *d_srcbuf, enum acgdatatype sendtype,
    int dstbufsize, void *d_dstbuf, enum acgdatatype recvtype,
    const struct acgcomm *comm, int tag, int *errcode, int warmup, cudaStream_t stream);
ACG_API int acghalo_exchange_cuda_end(
    struct acghalo *halo, struct acghaloexchange *haloexchange,
    int srcbufsize, const void *d_srcbuf, enum acgdatatype sendtype,
    int dstbufsize, void *d_dstbuf, enum acgdatatype recvtype,
    const struct acgcomm *comm, int tag, int *errcode, int warmup, cudaStream_t stream);
/* acg/halo.h:468 -- begin + end on the exchange's own stream, then synchronise */
ACG_API int acghalo_exchange_cuda(
    struct acghalo *halo, struct acghaloexchange *haloexchange,
    int srcbufsize, const void *d_srcbuf, enum acgdatatype sendtype,
    int dstbufsize, void *d_dstbuf, enum acgdatatype recvtype,
    const struct acgcomm *comm, int tag, int *errcode, int warmup);

#ifdef __cplusplus
}
#endif
#endif
