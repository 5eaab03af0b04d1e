/*
 * cuda_mock.c -- TEST INFRASTRUCTURE ONLY.  A stand-in for the handful of CUDA
 * runtime and NCCL entry points libacgb200's host code calls, so that the host
 * logic of the solver (set-up, iteration control, graph capture and replay,
 * polling, reports) can be exercised by the CPU test-suite on a machine
 * without a GPU.  "Device memory" is host memory, poisoned with NaN patterns
 * on allocation; streams and events are dummies (everything is synchronous);
 * a captured graph is a recorded list of the simulated kernel launches of
 * kernels_sim.c, replayed by cudaGraphLaunch.
 *
 * Built only into tests/hostsim/libacgb200_hostsim.so by tests/hostsim/Makefile.
 * Nothing under acg_b200/ refers to it: the product library links the real
 * CUDA runtime and fails loudly without a device.
 */
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hostsim.h"

struct simgraph *hostsim_capturing = NULL;

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double) ts.tv_sec + 1e-6 * (double) ts.tv_nsec;
}

/* ---- memory ---------------------------------------------------------------- */
cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xFF, n);                 /* NaN / -1: reading uninitialised "device" memory shows */
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *p, size_t n, unsigned int f) { (void) p; (void) n; (void) f; return cudaSuccess; }
cudaError_t cudaHostUnregister(void *p) { (void) p; return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind k) { (void) k; if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind k, cudaStream_t st)
{
    (void) k; (void) st;
    if (hostsim_capturing) return cudaErrorStreamCaptureUnsupported;     /* the iteration bodies contain kernels only */
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st)
{
    (void) st;
    if (hostsim_capturing) return cudaErrorStreamCaptureUnsupported;
    if (n) memset(d, v, n);
    return cudaSuccess;
}

/* ---- device, streams, events ----------------------------------------------- */
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned int f, int p)
{
    (void) f; (void) p;
    *s = (cudaStream_t) malloc(8);
    return *s ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void) s; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int f) { (void) s; (void) e; (void) f; return cudaSuccess; }

struct simevent { double t; };
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) calloc(1, sizeof(struct simevent)); return *e ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int f) { (void) f; return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void) s; ((struct simevent *) e)->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void) e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = (float) (((struct simevent *) b)->t - ((struct simevent *) a)->t);
    return cudaSuccess;
}

/* ---- graphs: a recorded list of simulated launches --------------------------- */
cudaError_t cudaStreamBeginCapture(cudaStream_t s, enum cudaStreamCaptureMode m)
{
    (void) s; (void) m;
    if (hostsim_capturing) return cudaErrorIllegalState;
    hostsim_capturing = calloc(1, sizeof(*hostsim_capturing));
    return hostsim_capturing ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g)
{
    (void) s;
    *g = (cudaGraph_t) hostsim_capturing;
    hostsim_capturing = NULL;
    return *g ? cudaSuccess : cudaErrorIllegalState;
}
static void graph_free(struct simgraph *g)
{
    if (!g) return;
    for (int i = 0; i < g->n; i++) free(g->ops[i].args);
    free(g->ops); free(g);
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { graph_free((struct simgraph *) g); return cudaSuccess; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long flags)
{
    (void) flags;
    const struct simgraph *src = (const struct simgraph *) g;
    struct simgraph *c = calloc(1, sizeof(*c));
    if (!c) return cudaErrorMemoryAllocation;
    c->ops = calloc((size_t) (src->n > 0 ? src->n : 1), sizeof(*c->ops));
    c->n = c->cap = src->n;
    for (int i = 0; i < src->n; i++) {
        c->ops[i] = src->ops[i];
        c->ops[i].args = malloc(src->ops[i].size);
        memcpy(c->ops[i].args, src->ops[i].args, src->ops[i].size);
    }
    *e = (cudaGraphExec_t) c;
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { graph_free((struct simgraph *) e); return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s)
{
    (void) s;
    const struct simgraph *g = (const struct simgraph *) e;
    for (int i = 0; i < g->n; i++) g->ops[i].fn(g->ops[i].args);
    return cudaSuccess;
}

int hostsim_run_or_record(void (*fn)(void *), const void *args, size_t size)
{
    if (!hostsim_capturing) { fn((void *) args); return 0; }
    struct simgraph *g = hostsim_capturing;
    if (g->n == g->cap) {
        g->cap = g->cap ? 2 * g->cap : 16;
        g->ops = realloc(g->ops, (size_t) g->cap * sizeof(*g->ops));
        if (!g->ops) return (int) cudaErrorMemoryAllocation;
    }
    g->ops[g->n].fn = fn; g->ops[g->n].size = size;
    g->ops[g->n].args = malloc(size);
    memcpy(g->ops[g->n].args, args, size);
    g->n++;
    return 0;
}

/* ---- inter-process pieces: not simulated (one process, null communicator) ---- */
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { (void) h; (void) p; return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int f) { (void) p; (void) h; (void) f; return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *p) { (void) p; return cudaErrorNotSupported; }

ncclResult_t ncclAllReduce(const void *s, void *r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) r; (void) n; (void) t; (void) o; (void) c; (void) st; return ncclInvalidUsage; }
ncclResult_t ncclAllGather(const void *s, void *r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) r; (void) n; (void) t; (void) c; (void) st; return ncclInvalidUsage; }
ncclResult_t ncclSend(const void *s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) n; (void) t; (void) peer; (void) c; (void) st; return ncclInvalidUsage; }
ncclResult_t ncclRecv(void *r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) r; (void) n; (void) t; (void) peer; (void) c; (void) st; return ncclInvalidUsage; }
ncclResult_t ncclGroupStart(void) { return ncclSuccess; }
ncclResult_t ncclGroupEnd(void) { return ncclSuccess; }
ncclResult_t ncclCommDestroy(ncclComm_t c) { (void) c; return ncclSuccess; }
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) { memset(id, 0, sizeof(*id)); return ncclInvalidUsage; }
ncclResult_t ncclCommInitRank(ncclComm_t *c, int n, ncclUniqueId id, int r) { (void) c; (void) n; (void) id; (void) r; return ncclInvalidUsage; }
ncclResult_t ncclCommCount(const ncclComm_t c, int *n) { (void) c; *n = 1; return ncclInvalidUsage; }
ncclResult_t ncclCommUserRank(const ncclComm_t c, int *r) { (void) c; *r = 0; return ncclInvalidUsage; }
ncclResult_t ncclCommSplit(ncclComm_t c, int color, int key, ncclComm_t *n, ncclConfig_t *cfg)
{ (void) c; (void) color; (void) key; (void) n; (void) cfg; return ncclInvalidUsage; }

/* ---- error strings ------------------------------------------------------------ */
const char *cudaGetErrorString(cudaError_t e) { (void) e; return "host-simulation stand-in: no CUDA error strings"; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char *ncclGetErrorString(ncclResult_t r) { (void) r; return "host-simulation stand-in: no NCCL error strings"; }
