/*
 * mpishim.c -- multi-process implementation of compat/mpi/mpi.h (see there for scope).
 *
 * One process per rank on one node.  Every pair of ranks is connected by a
 * Unix-domain stream socket (abstract namespace, full mesh set up in MPI_Init).
 * A message on the wire is a 24-byte header {magic, tag, context, bytes} followed
 * by the payload.  All sockets are non-blocking and there is ONE progress engine,
 * progress(): it polls every peer, appends what has arrived to per-peer queues of
 * complete messages, and writes out what is pending.  Sends are buffered (the
 * payload is copied into the peer's output queue), so a send never depends on the
 * receiver having posted a receive; a blocking call runs progress() until its own
 * condition holds (data written / matching message present), which is what keeps
 * head-to-head exchanges and "root scatters while ranks still send" free of deadlock.
 * Messages between two ranks are matched in order (non-overtaking) by tag and
 * communicator context.  Collectives are linear algorithms over point-to-point with
 * reserved negative tags; reductions combine in rank order, so every run gives the
 * same bits.
 *
 * Launch: compat/mpi/acgb200-mpirun (ACGB200_MPI_RANK / _SIZE / _JOB) or any
 * launcher exporting RANK, WORLD_SIZE and MASTER_PORT (torchrun --no-python).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <stddef.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#define MAGIC 0x4d504973u
#define TAG_BARRIER (-1001)
#define TAG_BCAST (-1002)
#define TAG_REDUCE (-1003)
#define TAG_GATHER (-1004)
#define TAG_SCATTER (-1005)
#define TAG_EXSCAN (-1006)

struct hdr { uint32_t magic; int32_t tag; int32_t ctx; int32_t pad; uint64_t bytes; };

struct msg { struct msg *next; int tag, ctx; size_t bytes; unsigned char *payload; };

struct peer {
    int fd;
    /* input: raw bytes not yet parsed, then complete messages in arrival order */
    unsigned char *in; size_t inlen, incap;
    struct msg *qhead, *qtail;
    /* output: pending bytes; `sent`/`queued` count bytes ever written / ever queued */
    unsigned char *out; size_t outlen, outoff, outcap;
    unsigned long long sent, queued;
};

struct comm { int used, ctx, size, rank, self; };

enum { REQ_FREE = 0, REQ_SEND, REQ_RECV };
struct req {
    int kind, persistent, active, done;
    void *buf; int count; MPI_Datatype t; int peer, tag; MPI_Comm comm;
    unsigned long long sendmark;      /* REQ_SEND: complete when peer.sent >= sendmark */
    MPI_Status st;
};

static int g_init, g_final, g_size = 1, g_rank;
static struct peer *g_peer;           /* [g_size]; g_peer[g_rank] is the loop-back queue */
static struct comm g_comm[64];
static int g_nextctx = 2;
static struct req *g_req;
static int g_nreq;
static char g_job[64];

static void die(const char *what)
{
    fprintf(stderr, "acgb200 mpishim (rank %d): %s: %s\n", g_rank, what, strerror(errno));
    _exit(86);
}

static struct comm *getcomm(MPI_Comm c) { return (c > 0 && c < 64 && g_comm[c].used) ? &g_comm[c] : NULL; }
static int world_of(const struct comm *c, int r) { return c->self ? g_rank : r; }

/* ---- progress engine -------------------------------------------------------------- */

static void q_push(struct peer *p, struct msg *m)
{
    m->next = NULL;
    if (p->qtail) p->qtail->next = m; else p->qhead = m;
    p->qtail = m;
}

static void parse_input(struct peer *p)
{
    size_t off = 0;
    while (p->inlen - off >= sizeof(struct hdr)) {
        struct hdr h;
        memcpy(&h, p->in + off, sizeof(h));
        if (h.magic != MAGIC) { errno = EPROTO; die("corrupt message stream"); }
        if (p->inlen - off - sizeof(h) < h.bytes) break;
        struct msg *m = malloc(sizeof(*m));
        if (!m) die("malloc");
        m->tag = h.tag; m->ctx = h.ctx; m->bytes = (size_t) h.bytes;
        m->payload = malloc(m->bytes ? m->bytes : 1);
        if (!m->payload) die("malloc");
        memcpy(m->payload, p->in + off + sizeof(h), m->bytes);
        q_push(p, m);
        off += sizeof(h) + m->bytes;
    }
    if (off) { memmove(p->in, p->in + off, p->inlen - off); p->inlen -= off; }
}

static void pump_in(struct peer *p)
{
    for (;;) {
        if (p->incap - p->inlen < (1u << 16)) {
            size_t nc = p->incap ? 2 * p->incap : (1u << 18);
            unsigned char *n = realloc(p->in, nc);
            if (!n) die("realloc");
            p->in = n; p->incap = nc;
        }
        ssize_t k = read(p->fd, p->in + p->inlen, p->incap - p->inlen);
        if (k > 0) { p->inlen += (size_t) k; if ((size_t) k < (1u << 16)) break; continue; }
        if (k == 0 || (k < 0 && errno == ECONNRESET && g_final)) {
            /* end of stream: fine once this rank is finalizing (the peer got everything it needed and left) */
            if (!g_final) { errno = ECONNRESET; die("a peer went away"); }
            close(p->fd); p->fd = -1;
            break;
        }
        if (errno == EAGAIN || errno == EWOULDBLOCK) break;
        if (errno == EINTR) continue;
        die("read");
    }
    parse_input(p);
}

static void pump_out(struct peer *p)
{
    if (p->fd < 0) { p->outoff = p->outlen = 0; return; }
    while (p->outoff < p->outlen) {
        ssize_t k = send(p->fd, p->out + p->outoff, p->outlen - p->outoff, MSG_NOSIGNAL);
        if (k > 0) { p->outoff += (size_t) k; p->sent += (unsigned long long) k; continue; }
        if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
        if (k < 0 && errno == EINTR) continue;
        if (k < 0 && g_final && (errno == EPIPE || errno == ECONNRESET)) { p->outoff = p->outlen = 0; return; }
        die("send");
    }
    if (p->outoff == p->outlen) p->outoff = p->outlen = 0;
}

/* one round: wait up to timeout_ms for any socket to become ready, move data both ways */
static void progress(int timeout_ms)
{
    if (g_size == 1) return;
    struct pollfd *pf = alloca((size_t) g_size * sizeof(*pf));
    int n = 0;
    for (int r = 0; r < g_size; r++) {
        if (r == g_rank || g_peer[r].fd < 0) continue;
        pf[n].fd = g_peer[r].fd; pf[n].events = POLLIN | (g_peer[r].outoff < g_peer[r].outlen ? POLLOUT : 0); pf[n].revents = 0;
        n++;
    }
    int k = poll(pf, (nfds_t) n, timeout_ms);
    if (k < 0 && errno != EINTR) die("poll");
    n = 0;
    for (int r = 0; r < g_size; r++) {
        if (r == g_rank || g_peer[r].fd < 0) continue;
        const short ev = pf[n++].revents;
        if (ev & POLLOUT) pump_out(&g_peer[r]);
        if (ev & (POLLIN | POLLHUP | POLLERR)) pump_in(&g_peer[r]);      /* may close the peer (finalize) */
    }
}

static void enqueue(int wdest, int tag, int ctx, const void *buf, size_t bytes, unsigned long long *mark)
{
    struct peer *p = &g_peer[wdest];
    if (wdest == g_rank) {
        struct msg *m = malloc(sizeof(*m));
        if (!m) die("malloc");
        m->tag = tag; m->ctx = ctx; m->bytes = bytes;
        m->payload = malloc(bytes ? bytes : 1);
        if (!m->payload) die("malloc");
        memcpy(m->payload, buf, bytes);
        q_push(p, m);
        if (mark) *mark = 0;
        return;
    }
    const size_t need = sizeof(struct hdr) + bytes;
    if (p->outcap - p->outlen < need) {
        if (p->outoff > 0) { memmove(p->out, p->out + p->outoff, p->outlen - p->outoff); p->outlen -= p->outoff; p->outoff = 0; }
        if (p->outcap - p->outlen < need) {
            size_t nc = p->outcap ? p->outcap : (1u << 16);
            while (nc - p->outlen < need) nc *= 2;
            unsigned char *n = realloc(p->out, nc);
            if (!n) die("realloc");
            p->out = n; p->outcap = nc;
        }
    }
    struct hdr h = { MAGIC, tag, ctx, 0, (uint64_t) bytes };
    memcpy(p->out + p->outlen, &h, sizeof(h));
    memcpy(p->out + p->outlen + sizeof(h), buf, bytes);
    p->outlen += need; p->queued += need;
    if (mark) *mark = p->queued;
    pump_out(p);
}

/* first queued message from world rank `wsrc` (or any, -1) matching tag / context; unlinks it */
static struct msg *match(int wsrc, int tag, int ctx, int *from)
{
    for (int r = (wsrc < 0 ? 0 : wsrc); r < (wsrc < 0 ? g_size : wsrc + 1); r++) {
        struct msg *prev = NULL;
        for (struct msg *m = g_peer[r].qhead; m; prev = m, m = m->next) {
            if (m->ctx != ctx || (tag != MPI_ANY_TAG && m->tag != tag) || (tag == MPI_ANY_TAG && m->tag < 0)) continue;
            if (prev) prev->next = m->next; else g_peer[r].qhead = m->next;
            if (g_peer[r].qtail == m) g_peer[r].qtail = prev;
            *from = r;
            return m;
        }
    }
    return NULL;
}

static int deliver(struct msg *m, int from, void *buf, size_t cap, MPI_Status *st)
{
    const int trunc = m->bytes > cap;
    memcpy(buf, m->payload, trunc ? cap : m->bytes);
    if (st) { st->MPI_SOURCE = from; st->MPI_TAG = m->tag; st->MPI_ERROR = trunc ? MPI_ERR_TRUNCATE : MPI_SUCCESS; st->nbytes_ = (long long) m->bytes; }
    free(m->payload); free(m);
    return trunc ? MPI_ERR_TRUNCATE : MPI_SUCCESS;
}

static int send_blocking(const void *buf, size_t bytes, int wdest, int tag, int ctx)
{
    unsigned long long mark = 0;
    enqueue(wdest, tag, ctx, buf, bytes, &mark);
    if (wdest != g_rank) while (g_peer[wdest].sent < mark) progress(1000);
    return MPI_SUCCESS;
}

static int recv_blocking(void *buf, size_t cap, int wsrc, int tag, int ctx, MPI_Status *st)
{
    for (;;) {
        int from = 0;
        struct msg *m = match(wsrc, tag, ctx, &from);
        if (m) return deliver(m, from, buf, cap, st);
        progress(1000);
    }
}

/* ---- environment ------------------------------------------------------------------ */

static void sockname(struct sockaddr_un *a, socklen_t *len, int rank)
{
    memset(a, 0, sizeof(*a));
    a->sun_family = AF_UNIX;
    /* abstract namespace: no file to clean up */
    const int n = snprintf(a->sun_path + 1, sizeof(a->sun_path) - 1, "acgb200mpi-%s-%d", g_job, rank);
    *len = (socklen_t) (offsetof(struct sockaddr_un, sun_path) + 1 + (size_t) n);
}

static void setnb(int fd)
{
    const int fl = fcntl(fd, F_GETFL, 0);
    if (fl < 0 || fcntl(fd, F_SETFL, fl | O_NONBLOCK) < 0) die("fcntl");
    int sz = 8 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
}

static void connect_mesh(void)
{
    struct sockaddr_un a; socklen_t al;
    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) die("socket");
    sockname(&a, &al, g_rank);
    if (bind(ls, (struct sockaddr *) &a, al) < 0) die("bind (is another job using the same ACGB200_MPI_JOB / MASTER_PORT?)");
    if (listen(ls, g_size) < 0) die("listen");
    for (int r = 0; r < g_rank; r++) {                 /* lower ranks are (or soon will be) listening */
        const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
        if (fd < 0) die("socket");
        sockname(&a, &al, r);
        int tries = 0;
        while (connect(fd, (struct sockaddr *) &a, al) < 0) {
            if ((errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN) || ++tries > 12000) die("connect to a lower rank (did it start?)");
            struct timespec ts = { 0, 10 * 1000 * 1000 };
            nanosleep(&ts, NULL);
        }
        int32_t me = g_rank;
        if (write(fd, &me, sizeof(me)) != (ssize_t) sizeof(me)) die("handshake");
        g_peer[r].fd = fd;
    }
    for (int k = g_rank + 1; k < g_size; k++) {
        const int fd = accept(ls, NULL, NULL);
        if (fd < 0) { if (errno == EINTR) { k--; continue; } die("accept"); }
        int32_t who = -1;
        size_t got = 0;
        while (got < sizeof(who)) {
            ssize_t n = read(fd, (char *) &who + got, sizeof(who) - got);
            if (n <= 0) { if (n < 0 && errno == EINTR) continue; die("handshake"); }
            got += (size_t) n;
        }
        if (who <= g_rank || who >= g_size || g_peer[who].fd >= 0) { errno = EPROTO; die("handshake: unexpected rank"); }
        g_peer[who].fd = fd;
    }
    close(ls);
    for (int r = 0; r < g_size; r++) if (r != g_rank) setnb(g_peer[r].fd);
}

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{
    (void) argc; (void) argv;
    if (provided) *provided = required < MPI_THREAD_FUNNELED ? required : MPI_THREAD_FUNNELED;
    if (g_init) return MPI_SUCCESS;
    const char *r = getenv("ACGB200_MPI_RANK"), *s = getenv("ACGB200_MPI_SIZE"), *j = getenv("ACGB200_MPI_JOB");
    if (!r || !s) { r = getenv("RANK"); s = getenv("WORLD_SIZE"); j = getenv("MASTER_PORT"); }
    if (r && s && atoi(s) > 1) {
        g_rank = atoi(r); g_size = atoi(s);
        snprintf(g_job, sizeof(g_job), "%s", j ? j : "0");
        if (g_rank < 0 || g_rank >= g_size) { fprintf(stderr, "acgb200 mpishim: bad rank %d of %d\n", g_rank, g_size); _exit(86); }
    }
    g_peer = calloc((size_t) g_size, sizeof(*g_peer));
    if (!g_peer) die("calloc");
    for (int i = 0; i < g_size; i++) g_peer[i].fd = -1;
    memset(g_comm, 0, sizeof(g_comm));
    g_comm[MPI_COMM_WORLD] = (struct comm) { 1, 0, g_size, g_rank, 0 };
    g_comm[MPI_COMM_SELF] = (struct comm) { 1, 1, 1, 0, 1 };
    if (g_size > 1) connect_mesh();
    g_init = 1;
    return MPI_SUCCESS;
}

int MPI_Init(int *argc, char ***argv) { return MPI_Init_thread(argc, argv, MPI_THREAD_SINGLE, NULL); }
int MPI_Initialized(int *flag) { *flag = g_init; return MPI_SUCCESS; }
int MPI_Query_thread(int *provided) { *provided = MPI_THREAD_FUNNELED; return MPI_SUCCESS; }

int MPI_Finalize(void)
{
    if (!g_init || g_final) return MPI_SUCCESS;
    g_final = 1;                            /* from here on a peer closing its end is not an error */
    MPI_Barrier(MPI_COMM_WORLD);            /* everything anybody still needs from me has been sent */
    for (int r = 0; r < g_size; r++) {
        if (r == g_rank || g_peer[r].fd < 0) continue;
        while (g_peer[r].fd >= 0 && g_peer[r].outoff < g_peer[r].outlen) progress(100);
    }
    for (int r = 0; r < g_size; r++) if (r != g_rank && g_peer[r].fd >= 0) { close(g_peer[r].fd); g_peer[r].fd = -1; }
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code)
{
    (void) comm;
    fflush(NULL);
    _exit(code ? code : 1);        /* the launcher ends the other ranks */
}

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int MPI_Get_processor_name(char *name, int *len)
{
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME - 1)) strcpy(name, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
    *len = (int) strlen(name);
    return MPI_SUCCESS;
}

int MPI_Get_library_version(char *version, int *len)
{
    strcpy(version, "acgb200 MPI shim (one node, Unix-domain sockets; compat/mpi/mpishim.c)");
    *len = (int) strlen(version);
    return MPI_SUCCESS;
}

int MPI_Error_string(int err, char *s, int *len)
{
    snprintf(s, MPI_MAX_ERROR_STRING, err == MPI_ERR_TRUNCATE ? "MPI shim: message truncated" : "MPI shim error %d", err);
    *len = (int) strlen(s);
    return MPI_SUCCESS;
}

/* ---- communicators, datatypes -------------------------------------------------------- */

int MPI_Comm_size(MPI_Comm comm, int *size) { struct comm *c = getcomm(comm); if (!c) return MPI_ERR_OTHER; *size = c->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { struct comm *c = getcomm(comm); if (!c) return MPI_ERR_OTHER; *rank = c->rank; return MPI_SUCCESS; }

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    for (int i = 3; i < 64; i++) {
        if (g_comm[i].used) continue;
        g_comm[i] = *c;
        g_comm[i].ctx = g_nextctx++;       /* collective: every rank draws the same number */
        *out = i;
        return MPI_SUCCESS;
    }
    return MPI_ERR_OTHER;
}

int MPI_Comm_free(MPI_Comm *comm)
{
    if (*comm > MPI_COMM_SELF && *comm < 64) g_comm[*comm].used = 0;
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

/* all ranks share the node */
int MPI_Comm_split_type(MPI_Comm comm, int type, int key, MPI_Info info, MPI_Comm *out)
{
    (void) type; (void) key; (void) info;
    return MPI_Comm_dup(comm, out);
}

int MPI_Type_size(MPI_Datatype t, int *size) { *size = ACGB200_MPI_SIZEOF(t); return MPI_SUCCESS; }
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newtype) { *newtype = 0x1000 + count * ACGB200_MPI_SIZEOF(old); return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype *t) { (void) t; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *t) { *t = MPI_DATATYPE_NULL; return MPI_SUCCESS; }

/* ---- point to point ------------------------------------------------------------------- */

static size_t nbytes(int count, MPI_Datatype t) { return (size_t) (count > 0 ? count : 0) * (size_t) ACGB200_MPI_SIZEOF(t); }

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c || dest < 0 || dest >= c->size || tag < 0) return MPI_ERR_OTHER;
    return send_blocking(buf, nbytes(count, t), world_of(c, dest), tag, c->ctx);
}

int MPI_Recv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Status *status)
{
    struct comm *c = getcomm(comm);
    if (!c || source >= c->size || (source < 0 && source != MPI_ANY_SOURCE)) return MPI_ERR_OTHER;
    return recv_blocking(buf, nbytes(count, t), source < 0 ? -1 : world_of(c, source), tag, c->ctx, status);
}

static int newreq(void)
{
    for (int i = 1; i < g_nreq; i++) if (g_req[i].kind == REQ_FREE) return i;
    const int n = g_nreq ? 2 * g_nreq : 64;
    struct req *r = realloc(g_req, (size_t) n * sizeof(*r));
    if (!r) die("realloc");
    memset(r + g_nreq, 0, (size_t) (n - g_nreq) * sizeof(*r));
    g_req = r;
    const int i = g_nreq ? g_nreq : 1;      /* 0 is MPI_REQUEST_NULL */
    g_nreq = n;
    return i;
}

static int mkreq(int kind, int persistent, void *buf, int count, MPI_Datatype t, int peer, int tag, MPI_Comm comm, MPI_Request *out)
{
    struct comm *c = getcomm(comm);
    if (!c || peer >= c->size) return MPI_ERR_OTHER;
    const int i = newreq();
    struct req *r = &g_req[i];
    memset(r, 0, sizeof(*r));
    r->kind = kind; r->persistent = persistent; r->buf = buf; r->count = count; r->t = t; r->peer = peer; r->tag = tag; r->comm = comm;
    *out = i;
    return MPI_SUCCESS;
}

static void start(struct req *r)
{
    struct comm *c = getcomm(r->comm);
    r->active = 1; r->done = 0;
    if (r->kind == REQ_SEND) enqueue(world_of(c, r->peer), r->tag, c->ctx, r->buf, nbytes(r->count, r->t), &r->sendmark);
}

/* one completion attempt; returns 1 when the request is complete */
static int try_complete(struct req *r)
{
    if (!r->active || r->done) return 1;
    struct comm *c = getcomm(r->comm);
    if (r->kind == REQ_SEND) {
        const int w = world_of(c, r->peer);
        if (w == g_rank || g_peer[w].sent >= r->sendmark) r->done = 1;
    } else {
        int from = 0;
        struct msg *m = match(r->peer < 0 ? -1 : world_of(c, r->peer), r->tag, c->ctx, &from);
        if (m) { deliver(m, from, r->buf, nbytes(r->count, r->t), &r->st); r->done = 1; }
    }
    return r->done;
}

static void retire(MPI_Request *req, MPI_Status *status)
{
    struct req *r = &g_req[*req];
    if (status && r->kind == REQ_RECV) *status = r->st;
    r->active = 0;
    if (!r->persistent) { r->kind = REQ_FREE; *req = MPI_REQUEST_NULL; }
}

int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    int e = mkreq(REQ_SEND, 0, (void *) buf, count, t, dest, tag, comm, req);
    if (!e) start(&g_req[*req]);
    return e;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
    int e = mkreq(REQ_RECV, 0, buf, count, t, source, tag, comm, req);
    if (!e) start(&g_req[*req]);
    return e;
}

int MPI_Send_init(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{ return mkreq(REQ_SEND, 1, (void *) buf, count, t, dest, tag, comm, req); }

int MPI_Recv_init(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req)
{ return mkreq(REQ_RECV, 1, buf, count, t, source, tag, comm, req); }

int MPI_Start(MPI_Request *req) { if (*req <= 0 || *req >= g_nreq) return MPI_ERR_OTHER; start(&g_req[*req]); return MPI_SUCCESS; }
int MPI_Startall(int n, MPI_Request *reqs) { for (int i = 0; i < n; i++) { int e = MPI_Start(&reqs[i]); if (e) return e; } return MPI_SUCCESS; }

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
    if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    if (*req < 0 || *req >= g_nreq || g_req[*req].kind == REQ_FREE) return MPI_ERR_OTHER;
    while (!try_complete(&g_req[*req])) progress(1000);
    retire(req, status);
    return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *statuses)
{
    for (int i = 0; i < n; i++) { int e = MPI_Wait(&reqs[i], statuses ? &statuses[i] : MPI_STATUS_IGNORE); if (e) return e; }
    return MPI_SUCCESS;
}

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status)
{
    *flag = 1;
    if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    if (*req < 0 || *req >= g_nreq || g_req[*req].kind == REQ_FREE) return MPI_ERR_OTHER;
    progress(0);
    if (try_complete(&g_req[*req])) retire(req, status); else *flag = 0;
    return MPI_SUCCESS;
}

int MPI_Request_free(MPI_Request *req)
{
    if (*req > 0 && *req < g_nreq) g_req[*req].kind = REQ_FREE;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status *status, MPI_Datatype t, int *count)
{
    const int sz = ACGB200_MPI_SIZEOF(t);
    *count = sz > 0 ? (int) (status->nbytes_ / sz) : 0;
    return MPI_SUCCESS;
}

/* ---- collectives (linear; reductions combine in rank order) ---------------------------- */

int MPI_Barrier(MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    if (c->size == 1) return MPI_SUCCESS;
    char b = 0;
    if (c->rank == 0) {
        for (int r = 1; r < c->size; r++) recv_blocking(&b, 1, world_of(c, r), TAG_BARRIER, c->ctx, NULL);
        for (int r = 1; r < c->size; r++) send_blocking(&b, 1, world_of(c, r), TAG_BARRIER, c->ctx);
    } else {
        send_blocking(&b, 1, world_of(c, 0), TAG_BARRIER, c->ctx);
        recv_blocking(&b, 1, world_of(c, 0), TAG_BARRIER, c->ctx, NULL);
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c || root < 0 || root >= c->size) return MPI_ERR_OTHER;
    if (c->size == 1) return MPI_SUCCESS;
    const size_t nb = nbytes(count, t);
    if (c->rank == root) {
        for (int r = 0; r < c->size; r++) if (r != root) enqueue(world_of(c, r), TAG_BCAST, c->ctx, buf, nb, NULL);
        for (int r = 0; r < c->size; r++) if (r != root) while (g_peer[world_of(c, r)].outoff < g_peer[world_of(c, r)].outlen) progress(1000);
        return MPI_SUCCESS;
    }
    return recv_blocking(buf, nb, world_of(c, root), TAG_BCAST, c->ctx, NULL);
}

#define COMBINE(T) do { T *a = inout; const T *b = in; for (int i = 0; i < count; i++) { \
    if (op == MPI_SUM) a[i] = (T) (a[i] + b[i]); else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; } \
    else if (op == MPI_MIN) { if (b[i] < a[i]) a[i] = b[i]; } else if (op == MPI_LOR) a[i] = (T) (a[i] || b[i]); \
    else if (op == MPI_LAND) a[i] = (T) (a[i] && b[i]); else return MPI_ERR_OTHER; } } while (0)

/* inout = inout (op) in, `in` being the contribution of the higher rank */
static int combine(void *inout, const void *in, int count, MPI_Datatype t, MPI_Op op)
{
    switch (t) {
    case MPI_INT: case MPI_INT32_T: COMBINE(int32_t); break;
    case MPI_UNSIGNED: COMBINE(uint32_t); break;
    case MPI_INT64_T: case MPI_LONG: COMBINE(int64_t); break;
    case MPI_UINT64_T: COMBINE(uint64_t); break;
    case MPI_DOUBLE: COMBINE(double); break;
    case MPI_FLOAT: COMBINE(float); break;
    case MPI_C_BOOL: case MPI_CHAR: case MPI_BYTE: COMBINE(unsigned char); break;
    case MPI_2INT: {
        if (op != MPI_MAXLOC) return MPI_ERR_OTHER;
        int32_t *a = inout; const int32_t *b = in;
        for (int i = 0; i < count; i++)
            if (b[2 * i] > a[2 * i] || (b[2 * i] == a[2 * i] && b[2 * i + 1] < a[2 * i + 1])) { a[2 * i] = b[2 * i]; a[2 * i + 1] = b[2 * i + 1]; }
        break;
    }
    default: return MPI_ERR_OTHER;
    }
    return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c || root < 0 || root >= c->size) return MPI_ERR_OTHER;
    const size_t nb = nbytes(count, t);
    if (c->rank != root) return send_blocking(sendbuf, nb, world_of(c, root), TAG_REDUCE, c->ctx);
    /* rank order 0, 1, ..., size-1 whatever the root: acc = c_0 op c_1 op ... */
    unsigned char *acc = malloc(nb ? nb : 1), *tmp = malloc(nb ? nb : 1);
    if (!acc || !tmp) die("malloc");
    int e = MPI_SUCCESS;
    for (int r = 0; r < c->size && !e; r++) {
        unsigned char *dst = r == 0 ? acc : tmp;
        if (r == root) memcpy(dst, sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, nb);
        else recv_blocking(dst, nb, world_of(c, r), TAG_REDUCE, c->ctx, NULL);
        if (r > 0) e = combine(acc, tmp, count, t, op);
    }
    if (!e) memcpy(recvbuf, acc, nb);
    free(acc); free(tmp);
    return e;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    int e = MPI_Reduce(c->rank == 0 ? sendbuf : (sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf), recvbuf, count, t, op, 0, comm);
    if (e) return e;
    return MPI_Bcast(recvbuf, count, t, 0, comm);
}

int MPI_Exscan(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    const size_t nb = nbytes(count, t);
    unsigned char *mine = malloc(nb ? nb : 1), *pre = malloc(nb ? nb : 1);
    if (!mine || !pre) die("malloc");
    memcpy(mine, sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, nb);
    int e = MPI_SUCCESS;
    if (c->rank > 0) {
        recv_blocking(pre, nb, world_of(c, c->rank - 1), TAG_EXSCAN, c->ctx, NULL);     /* c_0 op ... op c_{rank-1} */
        memcpy(recvbuf, pre, nb);
        e = combine(pre, mine, count, t, op);
    } else memcpy(pre, mine, nb);                                                        /* rank 0's result is undefined */
    if (!e && c->rank + 1 < c->size) send_blocking(pre, nb, world_of(c, c->rank + 1), TAG_EXSCAN, c->ctx);
    free(mine); free(pre);
    return e;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, const int *recvcounts,
                const int *displs, MPI_Datatype rt, int root, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c || root < 0 || root >= c->size) return MPI_ERR_OTHER;
    if (c->rank != root) return send_blocking(sendbuf, nbytes(sendcount, st), world_of(c, root), TAG_GATHER, c->ctx);
    const size_t rs = (size_t) ACGB200_MPI_SIZEOF(rt);
    for (int r = 0; r < c->size; r++) {
        unsigned char *dst = (unsigned char *) recvbuf + (size_t) displs[r] * rs;
        if (r == root) { if (sendbuf != MPI_IN_PLACE) memcpy(dst, sendbuf, nbytes(sendcount, st)); }
        else recv_blocking(dst, nbytes(recvcounts[r], rt), world_of(c, r), TAG_GATHER, c->ctx, NULL);
    }
    return MPI_SUCCESS;
}

int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt,
               int root, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    int *cnt = NULL, *dsp = NULL;
    if (c->rank == root) {
        cnt = malloc((size_t) c->size * sizeof(int)); dsp = malloc((size_t) c->size * sizeof(int));
        if (!cnt || !dsp) die("malloc");
        for (int r = 0; r < c->size; r++) { cnt[r] = recvcount; dsp[r] = r * recvcount; }
    }
    int e = MPI_Gatherv(sendbuf, sendcount, st, recvbuf, cnt, dsp, rt, root, comm);
    free(cnt); free(dsp);
    return e;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt,
                  MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c) return MPI_ERR_OTHER;
    const void *sb = sendbuf == MPI_IN_PLACE
        ? (const void *) ((unsigned char *) recvbuf + (size_t) c->rank * nbytes(recvcount, rt)) : sendbuf;
    void *tmp = NULL;
    if (c->rank == 0 && sendbuf == MPI_IN_PLACE) { tmp = malloc(nbytes(recvcount, rt) + 1); if (!tmp) die("malloc"); memcpy(tmp, sb, nbytes(recvcount, rt)); sb = tmp; }
    int e = MPI_Gather(sb, sendbuf == MPI_IN_PLACE ? recvcount : sendcount, sendbuf == MPI_IN_PLACE ? rt : st, recvbuf, recvcount, rt, 0, comm);
    free(tmp);
    if (e) return e;
    return MPI_Bcast(recvbuf, recvcount * c->size, rt, 0, comm);
}

int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype st, void *recvbuf,
                 int recvcount, MPI_Datatype rt, int root, MPI_Comm comm)
{
    struct comm *c = getcomm(comm);
    if (!c || root < 0 || root >= c->size) return MPI_ERR_OTHER;
    if (c->rank != root) return recv_blocking(recvbuf, nbytes(recvcount, rt), world_of(c, root), TAG_SCATTER, c->ctx, NULL);
    const size_t ss = (size_t) ACGB200_MPI_SIZEOF(st);
    for (int r = 0; r < c->size; r++) {
        const unsigned char *src = (const unsigned char *) sendbuf + (size_t) displs[r] * ss;
        if (r == root) { if (recvbuf != MPI_IN_PLACE) memcpy(recvbuf, src, nbytes(sendcounts[r], st)); }
        else send_blocking(src, nbytes(sendcounts[r], st), world_of(c, r), TAG_SCATTER, c->ctx);
    }
    return MPI_SUCCESS;
}
