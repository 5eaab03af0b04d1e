/*
 * nccl_mock.c -- TEST INFRASTRUCTURE ONLY (see cuda_mock.c).  The few NCCL calls
 * libacgb200's host code makes, for several processes on one machine, over files in
 * /dev/shm: a collective is "every rank drops a blob named after the communicator,
 * the collective's sequence number and its rank, then reads everybody's"; a
 * point-to-point message is a file named after sender, receiver and a per-pair
 * sequence number.  Calls block (the simulated device is synchronous); sends are
 * buffered, and inside ncclGroupStart/End receives are deferred behind the sends, so
 * the exchange patterns of halo.c cannot deadlock.  Slow and simple -- it only has to
 * carry test problems.
 */
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "hostsim.h"

#define MAXR 64
#define TIMEOUT_S 120.0

struct ncclComm {
    char name[40];
    int nranks, rank;
    unsigned long seq;                       /* collectives issued */
    unsigned long sendseq[MAXR], recvseq[MAXR];
    int nsplit;
};

struct pending { void *buf; size_t bytes; int peer; struct ncclComm *c; };
static int group_depth = 0;
static struct pending deferred[4 * MAXR];
static int ndeferred = 0;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void nap(void) { struct timespec ts = { 0, 50000 }; nanosleep(&ts, NULL); }

/* files this process created and nobody is known to have consumed: removed at exit (by then
 * every rank is past its last NCCL call -- the tests end with a barrier of their own) */
static char (*mine)[128] = NULL;
static int nmine = 0, capmine = 0;

static void cleanup_files(void)
{
    for (int i = 0; i < nmine; i++) unlink(mine[i]);
}

static void remember(const char *path)
{
    if (nmine == capmine) {
        if (capmine == 0) atexit(cleanup_files);
        capmine = capmine ? 2 * capmine : 64;
        mine = realloc(mine, (size_t) capmine * sizeof(*mine));
        if (!mine) { capmine = nmine = 0; return; }
    }
    snprintf(mine[nmine++], 128, "%s", path);
}

static void forget(const char *path)
{
    for (int i = nmine - 1; i >= 0; i--)
        if (!strcmp(mine[i], path)) { memmove(mine[i], mine[nmine - 1], 128); nmine--; return; }
}

static int put_file(const char *path, const void *buf, size_t bytes)
{
    char tmp[160];
    snprintf(tmp, sizeof(tmp), "%s.tmp%d", path, (int) getpid());
    FILE *f = fopen(tmp, "wb");
    if (!f) return -1;
    if (bytes && fwrite(buf, 1, bytes, f) != bytes) { fclose(f); unlink(tmp); return -1; }
    fclose(f);
    remember(path);
    return rename(tmp, path);                /* atomic: readers never see a partial file */
}

static int get_file(const char *path, void *buf, size_t bytes, int unlink_after)
{
    const double t0 = now_s();
    for (;;) {
        FILE *f = fopen(path, "rb");
        if (f) {
            const size_t got = bytes ? fread(buf, 1, bytes, f) : 0;
            fclose(f);
            if (got != bytes) return -1;
            if (unlink_after) unlink(path);
            return 0;
        }
        if (now_s() - t0 > TIMEOUT_S) return -1;
        nap();
    }
}

static void coll_path(char *out, size_t n, const struct ncclComm *c, unsigned long seq, int rank)
{
    snprintf(out, n, "/dev/shm/%s_c%lu_r%d", c->name, seq, rank);
}

/* every rank contributes `bytes`; all[] receives the contributions in rank order */
static int exchange(struct ncclComm *c, const void *mine, size_t bytes, void *all)
{
    char path[128];
    const unsigned long seq = c->seq++;
    if (seq >= 2) { coll_path(path, sizeof(path), c, seq - 2, c->rank); unlink(path); forget(path); }   /* everybody has read it */
    coll_path(path, sizeof(path), c, seq, c->rank);
    if (put_file(path, mine, bytes)) return -1;
    for (int r = 0; r < c->nranks; r++) {
        coll_path(path, sizeof(path), c, seq, r);
        if (get_file(path, (char *) all + (size_t) r * bytes, bytes, 0)) return -1;
    }
    return 0;
}

ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
    memset(id, 0, sizeof(*id));
    snprintf(id->internal, 32, "acgb200nccl_%d_%ld", (int) getpid(), (long) (now_s() * 1e6) % 100000000L);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank)
{
    if (nranks < 1 || nranks > MAXR || rank < 0 || rank >= nranks) return ncclInvalidArgument;
    struct ncclComm *c = calloc(1, sizeof(*c));
    if (!c) return ncclSystemError;
    memcpy(c->name, id.internal, 31);
    c->nranks = nranks; c->rank = rank;
    *comm = c;
    int one = 1, all[MAXR];
    return exchange(c, &one, sizeof(one), all) ? ncclSystemError : ncclSuccess;      /* rendezvous */
}

ncclResult_t ncclCommSplit(ncclComm_t comm, int color, int key, ncclComm_t *newcomm, ncclConfig_t *cfg)
{
    (void) color; (void) key; (void) cfg;       /* the library splits with one colour and key = rank */
    struct ncclComm *c = calloc(1, sizeof(*c));
    if (!c) return ncclSystemError;
    snprintf(c->name, sizeof(c->name), "%.30s_s%d", comm->name, comm->nsplit++);
    c->nranks = comm->nranks; c->rank = comm->rank;
    *newcomm = c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c)
{
    if (!c) return ncclSuccess;
    /* the last two collectives' files may still be needed by slower ranks: leave them to /dev/shm
     * housekeeping of the test (tests remove acgb200nccl_* afterwards) */
    free(c);
    return ncclSuccess;
}

ncclResult_t ncclCommCount(const ncclComm_t c, int *n) { *n = c->nranks; return ncclSuccess; }
ncclResult_t ncclCommUserRank(const ncclComm_t c, int *r) { *r = c->rank; return ncclSuccess; }

static size_t type_size(ncclDataType_t t)
{
    return t == ncclDouble ? 8 : t == ncclInt ? 4 : t == ncclChar ? 1 : 0;
}

ncclResult_t ncclAllGather(const void *s, void *r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t st)
{
    (void) st;
    const size_t bytes = n * type_size(t);
    if (!type_size(t)) return ncclInvalidArgument;
    return exchange(c, s, bytes, r) ? ncclSystemError : ncclSuccess;
}

static ncclResult_t allreduce_now(const void *s, void *r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c)
{
    const size_t ts = type_size(t), bytes = n * ts;
    if (!ts || (t != ncclDouble && t != ncclInt) || (o != ncclSum && o != ncclMin)) return ncclInvalidArgument;
    char *all = malloc(bytes * (size_t) c->nranks + 1);
    if (!all) return ncclSystemError;
    if (exchange(c, s, bytes, all)) { free(all); return ncclSystemError; }
    for (size_t i = 0; i < n; i++) {
        if (t == ncclDouble) {
            double acc = ((double *) all)[i];
            for (int q = 1; q < c->nranks; q++) {
                const double v = ((double *) (all + (size_t) q * bytes))[i];
                acc = o == ncclSum ? acc + v : (v < acc ? v : acc);
            }
            ((double *) r)[i] = acc;
        } else {
            int acc = ((int *) all)[i];
            for (int q = 1; q < c->nranks; q++) {
                const int v = ((int *) (all + (size_t) q * bytes))[i];
                acc = o == ncclSum ? acc + v : (v < acc ? v : acc);
            }
            ((int *) r)[i] = acc;
        }
    }
    free(all);
    return ncclSuccess;
}

static void p2p_path(char *out, size_t n, const struct ncclComm *c, int src, int dst, unsigned long seq)
{
    snprintf(out, n, "/dev/shm/%s_p%d_%d_%lu", c->name, src, dst, seq);
}

static ncclResult_t send_now(const void *s, size_t n, ncclDataType_t t, int peer, ncclComm_t c)
{
    char path[128];
    p2p_path(path, sizeof(path), c, c->rank, peer, c->sendseq[peer]++);
    return put_file(path, s, n * type_size(t)) ? ncclSystemError : ncclSuccess;       /* buffered */
}

static ncclResult_t do_recv(void *r, size_t bytes, int peer, struct ncclComm *c)
{
    char path[128];
    p2p_path(path, sizeof(path), c, peer, c->rank, c->recvseq[peer]++);
    return get_file(path, r, bytes, 1) ? ncclSystemError : ncclSuccess;
}

static ncclResult_t recv_now(void *r, size_t n, ncclDataType_t t, int peer, ncclComm_t c)
{
    const size_t bytes = n * type_size(t);
    if (group_depth > 0) {
        if (ndeferred >= (int) (sizeof(deferred) / sizeof(deferred[0]))) return ncclInternalError;
        deferred[ndeferred].buf = r; deferred[ndeferred].bytes = bytes; deferred[ndeferred].peer = peer;
        deferred[ndeferred].c = c; ndeferred++;
        return ncclSuccess;
    }
    return do_recv(r, bytes, peer, c);
}

static ncclResult_t groupstart_now(void) { group_depth++; return ncclSuccess; }

static ncclResult_t groupend_now(void)
{
    if (--group_depth > 0) return ncclSuccess;
    ncclResult_t res = ncclSuccess;
    for (int i = 0; i < ndeferred; i++) {
        const ncclResult_t r = do_recv(deferred[i].buf, deferred[i].bytes, deferred[i].peer, deferred[i].c);
        if (r != ncclSuccess) res = r;
    }
    ndeferred = 0;
    return res;
}

/* The calls of the data path may sit inside a captured iteration (cudaStreamBeginCapture ...
 * cudaGraphLaunch): like the simulated kernels they are then recorded and replayed. */
struct nop { int kind; const void *s; void *r; size_t n; ncclDataType_t t; ncclRedOp_t o; int peer; ncclComm_t c; };
static ncclResult_t replay_status = ncclSuccess;

static void nop_exec(void *vp)
{
    const struct nop *a = vp;
    ncclResult_t r = ncclSuccess;
    switch (a->kind) {
    case 0: r = allreduce_now(a->s, a->r, a->n, a->t, a->o, a->c); break;
    case 1: r = send_now(a->s, a->n, a->t, a->peer, a->c); break;
    case 2: r = recv_now(a->r, a->n, a->t, a->peer, a->c); break;
    case 3: r = groupstart_now(); break;
    case 4: r = groupend_now(); break;
    }
    if (r != ncclSuccess) replay_status = r;
}

static ncclResult_t submit(const struct nop *a)
{
    replay_status = ncclSuccess;
    if (hostsim_run_or_record(nop_exec, a, sizeof(*a))) return ncclSystemError;
    return replay_status;
}

ncclResult_t ncclAllReduce(const void *s, void *r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c, cudaStream_t st)
{
    (void) st;
    /* NCCL discards empty collectives at enqueue (no rendezvous at all): mirror that, so that code
     * relying on a zero-count allreduce as a barrier fails here as it would on the real library */
    if (n == 0) return ncclSuccess;
    const struct nop a = { 0, s, r, n, t, o, 0, c };
    return submit(&a);
}
ncclResult_t ncclSend(const void *s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{
    (void) st;
    const struct nop a = { 1, s, NULL, n, t, ncclSum, peer, c };
    return submit(&a);
}
ncclResult_t ncclRecv(void *r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{
    (void) st;
    const struct nop a = { 2, NULL, r, n, t, ncclSum, peer, c };
    return submit(&a);
}
ncclResult_t ncclGroupStart(void) { const struct nop a = { 3, NULL, NULL, 0, ncclChar, ncclSum, 0, NULL }; return submit(&a); }
ncclResult_t ncclGroupEnd(void) { const struct nop a = { 4, NULL, NULL, 0, ncclChar, ncclSum, 0, NULL }; return submit(&a); }

const char *ncclGetErrorString(ncclResult_t r) { (void) r; return "host-simulation stand-in: no NCCL error strings"; }
