/* p2p.h -- host side of the peer-memory exchange (see p2p.c, internal.h) */
#ifndef ACGB200_P2P_H
#define ACGB200_P2P_H

#include "acgb200/comm.h"
#include "acgb200/halo.h"
#include "internal.h"

struct acgb200_p2p {
    int enabled;
    int nranks, rank;
    void *window;                       /* this rank's exported allocation */
    size_t window_bytes;
    void *peer_base[ACGB200_MAXR];      /* mapped windows (own window at [rank]) */
    struct acgb200_p2pdev h_desc;       /* host mirror */
    struct acgb200_p2pdev *d_desc;      /* device descriptor the kernels read */
    unsigned long long seq;             /* next unused sequence number */
};

int acgb200_p2p_init(struct acgb200_p2p *p, const struct acghalo *halo, int borderoff, int nborder,
                     const struct acgcomm *comm, cudaStream_t stream, int *errcode);
int acgb200_p2p_begin(struct acgb200_p2p *p, int maxits, cudaStream_t stream);
void acgb200_p2p_free(struct acgb200_p2p *p);
/* after a solve: *flag = 1 if a device-side wait for a peer gave up (ACGB200_P2P_TIMEOUT_MS, default 30 s) */
int acgb200_p2p_timed_out(struct acgb200_p2p *p, cudaStream_t stream, int *flag);
int acgb200_p2p_inverse_map(const struct acghalo *halo, int borderoff, int nborder,
                            const int *rdispl_at_recipient, int *bptr, int *bq, int *bdst);

#endif
