/*
 * acgb200/comm.h -- communicator abstraction (NCCL over NVLink 5 / NVSwitch).
 *
 * ABI counterpart of acg/comm.h:84-117 (enum acgcommtype, struct acgcomm)
 * and :135-266 (init / size / rank / barrier / allreduce).  Only the NCCL
 * back-end carries data in this build; the MPI member exists (when compiled
 * with ACG_HAVE_MPI, as the reference driver is) purely so that the struct
 * layout matches what cuda/acg-cuda.c allocates.
 */
#ifndef ACGB200_COMM_H
#define ACGB200_COMM_H

#include "acgb200/config.h"

#ifdef ACG_HAVE_MPI
#include <mpi.h>
#endif
#include <cuda_runtime_api.h>
#include <nccl.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(ACG_HAVE_MPI)
#define ACG_IN_PLACE MPI_IN_PLACE
#else
#define ACG_IN_PLACE ((void *) 1)        /* acg/comm.h:74 */
#endif

/* acg/comm.h:84-92 */
enum acgcommtype {
    acgcomm_null, acgcomm_mpi, acgcomm_nccl, acgcomm_rccl, acgcomm_nvshmem, acgcomm_rocshmem,
};
ACG_API const char *acgcommtypestr(enum acgcommtype commtype);

/* acg/comm.h:103-117 */
struct acgcomm {
    enum acgcommtype type;
#if defined(ACG_HAVE_MPI)
    MPI_Comm mpicomm;
#endif
    ncclComm_t ncclcomm;
};

/* acg/comm.h:180-183, :219-222 */
enum acgdatatype { ACG_DOUBLE };
enum acgop { ACG_SUM };

/* acg/comm.h:188-240 */
ACG_API const char *acgdatatypestr(enum acgdatatype datatype);
ACG_API int acgdatatype_size(enum acgdatatype datatype, int *size);
ACG_API ncclDataType_t acgdatatype_nccl(enum acgdatatype datatype);
ACG_API const char *acgopstr(enum acgop op);
ACG_API ncclRedOp_t acgop_nccl(enum acgop op);

/* acg/comm.h:135 -- wraps (does not own) an existing NCCL communicator */
ACG_API int acgcomm_init_nccl(struct acgcomm *comm, ncclComm_t ncclcomm, int *ncclerrcode);
#if defined(ACG_HAVE_MPI)
/* acg/comm.h:124 -- layout/bootstrap only; data-path calls on it return ACG_ERR_MPI_NOT_SUPPORTED */
ACG_API int acgcomm_init_mpi(struct acgcomm *comm, MPI_Comm mpicomm, int *mpierrcode);
#endif
/* acg/comm.h:155 */ ACG_API void acgcomm_free(struct acgcomm *comm);
/* acg/comm.h:161 */ ACG_API int acgcomm_size(const struct acgcomm *comm, int *commsize);
/* acg/comm.h:168 */ ACG_API int acgcomm_rank(const struct acgcomm *comm, int *rank);
/* acg/comm.h:251 -- zero-byte allreduce on the stream (acg/comm.c:331) */
ACG_API int acgcomm_barrier(cudaStream_t stream, const struct acgcomm *comm, int *errcode);
/* acg/comm.h:259 -- ncclAllReduce on the stream; src may be ACG_IN_PLACE (acg/comm.c:350-398) */
ACG_API int acgcomm_allreduce(const void *src, void *dst, int count, enum acgdatatype datatype,
                              enum acgop op, cudaStream_t stream, const struct acgcomm *comm, int *errcode);

#ifdef __cplusplus
}
#endif
#endif
