/*
 * comm.c -- communicator wrapper: NCCL over NVLink 5 / NVSwitch.
 *
 * Own implementation of the acgcomm_* entry points the solver path uses
 * (reference: acg/comm.c:76-398).  The communicator is borrowed, never
 * destroyed here (acg/comm.c:139-141).
 */
#include "acgb200/comm.h"
#include "acgb200/error.h"

#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stddef.h>

const char *acgcommtypestr(enum acgcommtype t)
{
    switch (t) {
    case acgcomm_null: return "null";
    case acgcomm_mpi: return "mpi";
    case acgcomm_nccl: return "nccl";
    case acgcomm_rccl: return "rccl";
    case acgcomm_nvshmem: return "nvshmem";
    case acgcomm_rocshmem: return "rocshmem";
    default: return "unknown";
    }
}

/* acg/comm.h:188-240 */
const char *acgdatatypestr(enum acgdatatype datatype) { return datatype == ACG_DOUBLE ? "double" : "unknown"; }
const char *acgopstr(enum acgop op) { return op == ACG_SUM ? "sum" : "unknown"; }
ncclDataType_t acgdatatype_nccl(enum acgdatatype datatype) { (void) datatype; return ncclDouble; }
ncclRedOp_t acgop_nccl(enum acgop op) { (void) op; return ncclSum; }

int acgdatatype_size(enum acgdatatype datatype, int *size)
{
    if (datatype != ACG_DOUBLE) return ACG_ERR_INVALID_VALUE;
    *size = (int) sizeof(double);
    return ACG_SUCCESS;
}

int acgcomm_init_nccl(struct acgcomm *comm, ncclComm_t ncclcomm, int *ncclerrcode)
{
    (void) ncclerrcode;
    comm->type = acgcomm_nccl;
#if defined(ACG_HAVE_MPI)
    comm->mpicomm = MPI_COMM_NULL;
#endif
    comm->ncclcomm = ncclcomm;
    return ACG_SUCCESS;
}

#if defined(ACG_HAVE_MPI)
int acgcomm_init_mpi(struct acgcomm *comm, MPI_Comm mpicomm, int *mpierrcode)
{
    (void) mpierrcode;
    comm->type = acgcomm_mpi;
    comm->mpicomm = mpicomm;
    comm->ncclcomm = NULL;
    return ACG_SUCCESS;
}
#endif

void acgcomm_free(struct acgcomm *comm)
{
    comm->type = acgcomm_null;
    comm->ncclcomm = NULL;
}

int acgcomm_size(const struct acgcomm *comm, int *commsize)
{
    if (!comm || comm->type == acgcomm_null) { *commsize = 1; return ACG_SUCCESS; }
    if (comm->type == acgcomm_nccl) {
        ncclResult_t r = ncclCommCount(comm->ncclcomm, commsize);
        return r == ncclSuccess ? ACG_SUCCESS : ACG_ERR_NCCL;
    }
#if defined(ACG_HAVE_MPI)
    if (comm->type == acgcomm_mpi) return MPI_Comm_size(comm->mpicomm, commsize) ? ACG_ERR_MPI : ACG_SUCCESS;
#endif
    return ACG_ERR_INVALID_VALUE;
}

int acgcomm_rank(const struct acgcomm *comm, int *rank)
{
    if (!comm || comm->type == acgcomm_null) { *rank = 0; return ACG_SUCCESS; }
    if (comm->type == acgcomm_nccl) {
        ncclResult_t r = ncclCommUserRank(comm->ncclcomm, rank);
        return r == ncclSuccess ? ACG_SUCCESS : ACG_ERR_NCCL;
    }
#if defined(ACG_HAVE_MPI)
    if (comm->type == acgcomm_mpi) return MPI_Comm_rank(comm->mpicomm, rank) ? ACG_ERR_MPI : ACG_SUCCESS;
#endif
    return ACG_ERR_INVALID_VALUE;
}

int acgcomm_allreduce(const void *src, void *dst, int count, enum acgdatatype datatype,
                      enum acgop op, cudaStream_t stream, const struct acgcomm *comm, int *errcode)
{
    if (datatype != ACG_DOUBLE || op != ACG_SUM) return ACG_ERR_INVALID_VALUE;
    if (!comm || comm->type == acgcomm_null) {
        if (src != ACG_IN_PLACE && src != dst) {
            cudaError_t e = cudaMemcpyAsync(dst, src, (size_t) count * sizeof(double), cudaMemcpyDeviceToDevice, stream);
            if (e) { if (errcode) *errcode = (int) e; return ACG_ERR_CUDA; }
        }
        return ACG_SUCCESS;
    }
    if (comm->type == acgcomm_nccl) {
        if (src == ACG_IN_PLACE) src = dst;
        ncclResult_t r = ncclAllReduce(src, dst, (size_t) count, ncclDouble, ncclSum, comm->ncclcomm, stream);
        if (r != ncclSuccess) { if (errcode) *errcode = (int) r; return ACG_ERR_NCCL; }
        return ACG_SUCCESS;
    }
    if (comm->type == acgcomm_mpi) {
        /* MPI carries no data in this build; a one-rank MPI communicator (the
         * reference driver's default --comm mpi on a single GPU) is a no-op */
        int size = 1;
        if (acgcomm_size(comm, &size) == ACG_SUCCESS && size == 1)
            return acgcomm_allreduce(src, dst, count, datatype, op, stream, NULL, errcode);
        return ACG_ERR_MPI_NOT_SUPPORTED;
    }
    if (comm->type == acgcomm_nvshmem) return ACG_ERR_NVSHMEM_NOT_SUPPORTED;
    return ACG_ERR_INVALID_VALUE;
}

/* one double per device for the barrier's allreduce (never read; zero + zero stays zero) */
#define BARRIER_MAXDEV 64
static double *barrier_scratch[BARRIER_MAXDEV];
static pthread_mutex_t barrier_lock = PTHREAD_MUTEX_INITIALIZER;

static double *barrier_buffer(int *errcode)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (!e && (dev < 0 || dev >= BARRIER_MAXDEV)) e = cudaErrorInvalidDevice;
    double *p = NULL;
    if (!e) {
        pthread_mutex_lock(&barrier_lock);
        if (!barrier_scratch[dev]) {
            e = cudaMalloc((void **) &barrier_scratch[dev], 2 * sizeof(double));
            if (!e) e = cudaMemset(barrier_scratch[dev], 0, 2 * sizeof(double));
            if (e) { cudaFree(barrier_scratch[dev]); barrier_scratch[dev] = NULL; }
        }
        p = barrier_scratch[dev];
        pthread_mutex_unlock(&barrier_lock);
    }
    if (e && errcode) *errcode = (int) e;
    return p;
}

int acgcomm_barrier(cudaStream_t stream, const struct acgcomm *comm, int *errcode)
{
    if (!comm || comm->type == acgcomm_null) return ACG_SUCCESS;
    if (comm->type == acgcomm_nccl) {
        /* The reference enqueues a zero-length allreduce here (acg/comm.c:331).  NCCL drops empty
         * collectives at enqueue without any rendezvous, so that orders nothing; a one-element
         * allreduce does: work enqueued on `stream` behind it starts only after every rank has
         * reached its own barrier call on its stream.  Callers that need the HOST ordered behind
         * all ranks synchronise the stream afterwards. */
        double *buf = barrier_buffer(errcode);
        if (!buf) return ACG_ERR_CUDA;
        ncclResult_t r = ncclAllReduce(buf, buf, 1, ncclDouble, ncclSum, comm->ncclcomm, stream);
        if (r != ncclSuccess) { if (errcode) *errcode = (int) r; return ACG_ERR_NCCL; }
        return ACG_SUCCESS;
    }
    if (comm->type == acgcomm_mpi) {
        int size = 1;
        if (acgcomm_size(comm, &size) == ACG_SUCCESS && size == 1) return ACG_SUCCESS;
        return ACG_ERR_MPI_NOT_SUPPORTED;
    }
    return ACG_ERR_INVALID_VALUE;
}
