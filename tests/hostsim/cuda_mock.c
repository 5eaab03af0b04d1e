/*
 * cuda_mock.c -- TEST INFRASTRUCTURE ONLY.  A stand-in for the handful of CUDA
 * runtime and NCCL entry points libacgb200's host code calls, so that the host
 * logic of the solver (set-up, iteration control, graph capture and replay,
 * polling, reports) can be exercised by the CPU test-suite on a machine
 * without a GPU.  "Device memory" is host memory, poisoned with NaN patterns
 * on allocation; streams and events are dummies (everything is synchronous);
 * a captured graph is a recorded list of the simulated kernel launches of
 * kernels_sim.c, replayed by cudaGraphLaunch.  Allocations are shared-memory
 * objects so that multi-process tests can export and map them like CUDA IPC;
 * the NCCL stand-in is nccl_mock.c.
 *
 * Built only into tests/hostsim/libacgb200_hostsim.so by tests/hostsim/Makefile.
 * Nothing under acg_b200/ refers to it: the product library links the real
 * CUDA runtime and fails loudly without a device.
 */
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "hostsim.h"

struct simgraph *hostsim_capturing = NULL;

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double) ts.tv_sec + 1e-6 * (double) ts.tv_nsec;
}

/* ---- memory ---------------------------------------------------------------- */
/* Every "device" allocation is a POSIX shared-memory object, so that any of them
 * can be exported with cudaIpcGetMemHandle and mapped by another process of a
 * multi-rank test: the peer-memory exchange then really runs between processes. */
struct simalloc { void *ptr; size_t size; char name[48]; int mapped_peer; };
static struct simalloc allocs[1024];
static int nallocs = 0;
static pthread_mutex_t alloc_lock = PTHREAD_MUTEX_INITIALIZER;
static unsigned long alloc_counter = 0;

static void *shm_map(const char *name, size_t n, int create)
{
    const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return NULL;
    if (create && ftruncate(fd, (off_t) n) != 0) { close(fd); shm_unlink(name); return NULL; }
    void *p = mmap(NULL, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    return p == MAP_FAILED ? NULL : p;
}

static void cleanup_allocs(void)
{
    for (int i = 0; i < nallocs; i++) if (allocs[i].ptr && !allocs[i].mapped_peer) shm_unlink(allocs[i].name);
}

/* fault injection for the clean-up paths: HOSTSIM_FAIL_MALLOC_AT=k makes the k-th device allocation
 * fail, HOSTSIM_FAIL_CALL_AT=k the k-th call of any runtime entry point that can fail */
static long malloc_calls = 0, api_calls = 0;

static int failpoint(void)
{
    const char *at = getenv("HOSTSIM_FAIL_CALL_AT");
    return at && ++api_calls == atol(at);
}
#define FAILPOINT(err) do { if (failpoint()) return (err); } while (0)

cudaError_t cudaMalloc(void **p, size_t n)
{ FAILPOINT(cudaErrorUnknown);
    const size_t size = n ? n : 1;
    const char *failat = getenv("HOSTSIM_FAIL_MALLOC_AT");
    if (failat && ++malloc_calls == atol(failat)) { *p = NULL; return cudaErrorMemoryAllocation; }
    pthread_mutex_lock(&alloc_lock);
    if (nallocs == 0) atexit(cleanup_allocs);
    int slot = -1;
    for (int i = 0; i < nallocs; i++) if (!allocs[i].ptr) { slot = i; break; }
    if (slot < 0 && nallocs < 1024) slot = nallocs++;
    if (slot < 0) { pthread_mutex_unlock(&alloc_lock); return cudaErrorMemoryAllocation; }
    snprintf(allocs[slot].name, sizeof(allocs[slot].name), "/acgb200sim_%d_%lu", (int) getpid(), alloc_counter++);
    void *q = shm_map(allocs[slot].name, size, 1);
    if (!q) { pthread_mutex_unlock(&alloc_lock); return cudaErrorMemoryAllocation; }
    allocs[slot].ptr = q; allocs[slot].size = size; allocs[slot].mapped_peer = 0;
    pthread_mutex_unlock(&alloc_lock);
    memset(q, 0xFF, size);               /* NaN / -1: reading uninitialised "device" memory shows */
    *p = q;
    return cudaSuccess;
}

cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    pthread_mutex_lock(&alloc_lock);
    for (int i = 0; i < nallocs; i++) {
        if (allocs[i].ptr == p && !allocs[i].mapped_peer) {
            munmap(p, allocs[i].size);
            shm_unlink(allocs[i].name);
            allocs[i].ptr = NULL;
            pthread_mutex_unlock(&alloc_lock);
            return cudaSuccess;
        }
    }
    pthread_mutex_unlock(&alloc_lock);
    return cudaErrorInvalidValue;
}

/* the handle carries the name and the size of the object */
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    memset(h, 0, sizeof(*h));
    pthread_mutex_lock(&alloc_lock);
    for (int i = 0; i < nallocs; i++) {
        if (allocs[i].ptr == p && !allocs[i].mapped_peer) {
            memcpy(h->reserved, allocs[i].name, sizeof(allocs[i].name));
            memcpy(h->reserved + 48, &allocs[i].size, sizeof(size_t));
            pthread_mutex_unlock(&alloc_lock);
            return cudaSuccess;
        }
    }
    pthread_mutex_unlock(&alloc_lock);
    return cudaErrorInvalidValue;
}

cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int f)
{
    (void) f;
    char name[48];
    size_t size = 0;
    memcpy(name, h.reserved, sizeof(name)); name[47] = 0;
    memcpy(&size, h.reserved + 48, sizeof(size_t));
    void *q = shm_map(name, size, 0);
    if (!q) return cudaErrorInvalidValue;
    pthread_mutex_lock(&alloc_lock);
    int slot = -1;
    for (int i = 0; i < nallocs; i++) if (!allocs[i].ptr) { slot = i; break; }
    if (slot < 0 && nallocs < 1024) slot = nallocs++;
    if (slot >= 0) { allocs[slot].ptr = q; allocs[slot].size = size; allocs[slot].mapped_peer = 1; allocs[slot].name[0] = 0; }
    pthread_mutex_unlock(&alloc_lock);
    *p = q;
    return cudaSuccess;
}

cudaError_t cudaIpcCloseMemHandle(void *p)
{
    pthread_mutex_lock(&alloc_lock);
    for (int i = 0; i < nallocs; i++) {
        if (allocs[i].ptr == p && allocs[i].mapped_peer) {
            munmap(p, allocs[i].size);
            allocs[i].ptr = NULL;
            pthread_mutex_unlock(&alloc_lock);
            return cudaSuccess;
        }
    }
    pthread_mutex_unlock(&alloc_lock);
    return cudaErrorInvalidValue;
}

cudaError_t cudaMallocHost(void **p, size_t n) { FAILPOINT(cudaErrorUnknown); *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *p, size_t n, unsigned int f) { (void) p; (void) n; (void) f; return cudaSuccess; }
cudaError_t cudaHostUnregister(void *p) { (void) p; return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind k) { FAILPOINT(cudaErrorUnknown); (void) k; if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind k, cudaStream_t st)
{ FAILPOINT(cudaErrorUnknown);
    (void) k; (void) st;
    if (hostsim_capturing) return cudaErrorStreamCaptureUnsupported;     /* the iteration bodies contain kernels only */
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) { FAILPOINT(cudaErrorUnknown); if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st)
{ FAILPOINT(cudaErrorUnknown);
    (void) st;
    if (hostsim_capturing) return cudaErrorStreamCaptureUnsupported;
    if (n) memset(d, v, n);
    return cudaSuccess;
}

/* ---- device, streams, events ----------------------------------------------- */
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { FAILPOINT(cudaErrorUnknown); *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned int f, int p)
{ FAILPOINT(cudaErrorUnknown);
    (void) f; (void) p;
    *s = (cudaStream_t) malloc(8);
    return *s ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { FAILPOINT(cudaErrorUnknown); (void) s; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int f) { (void) s; (void) e; (void) f; return cudaSuccess; }

struct simevent { double t; };
cudaError_t cudaEventCreate(cudaEvent_t *e) { FAILPOINT(cudaErrorUnknown); *e = (cudaEvent_t) calloc(1, sizeof(struct simevent)); return *e ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int f) { (void) f; return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { FAILPOINT(cudaErrorUnknown); (void) s; ((struct simevent *) e)->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void) e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = (float) (((struct simevent *) b)->t - ((struct simevent *) a)->t);
    return cudaSuccess;
}

/* ---- graphs: a recorded list of simulated launches --------------------------- */
cudaError_t cudaStreamBeginCapture(cudaStream_t s, enum cudaStreamCaptureMode m)
{ FAILPOINT(cudaErrorUnknown);
    (void) s; (void) m;
    if (hostsim_capturing) return cudaErrorIllegalState;
    hostsim_capturing = calloc(1, sizeof(*hostsim_capturing));
    return hostsim_capturing ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g)
{ FAILPOINT(cudaErrorUnknown);
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
{ FAILPOINT(cudaErrorUnknown);
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
{ FAILPOINT(cudaErrorUnknown);
    (void) s;
    const struct simgraph *g = (const struct simgraph *) e;
    for (int i = 0; i < g->n; i++) g->ops[i].fn(g->ops[i].args);
    return cudaSuccess;
}

int hostsim_run_or_record(void (*fn)(void *), const void *args, size_t size)
{
    {   /* HOSTSIM_FAIL_LAUNCH_AT=k: the k-th kernel launch (or NCCL call) reports an error */
        static long launches = 0;
        const char *at = getenv("HOSTSIM_FAIL_LAUNCH_AT");
        if (at && ++launches == atol(at)) return (int) cudaErrorLaunchFailure;
    }
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

/* ---- error strings ------------------------------------------------------------ */
const char *cudaGetErrorString(cudaError_t e) { (void) e; return "host-simulation stand-in: no CUDA error strings"; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
