/*
 * ext.c -- small helpers for bindings and multi-process bootstrap
 * (include/acgb200/ext.h).
 */
#include "acgb200/error.h"
#include "acgb200/ext.h"

#include <cuda_runtime_api.h>
#include <string.h>

size_t acgb200_sizeof(const char *name)
{
    if (!strcmp(name, "acgsolvercuda")) return sizeof(struct acgsolvercuda);
    if (!strcmp(name, "acgsymcsrmatrix")) return sizeof(struct acgsymcsrmatrix);
    if (!strcmp(name, "acgvector")) return sizeof(struct acgvector);
    if (!strcmp(name, "acgcomm")) return sizeof(struct acgcomm);
    if (!strcmp(name, "acghalo")) return sizeof(struct acghalo);
    if (!strcmp(name, "acghaloexchange")) return sizeof(struct acghaloexchange);
    if (!strcmp(name, "acggraph")) return sizeof(struct acggraph);
    return 0;
}

int acgb200_have_mpi(void)
{
#ifdef ACG_HAVE_MPI
    return 1;
#else
    return 0;
#endif
}

int acgb200_host_register(void *ptr, size_t bytes)
{
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? ACG_SUCCESS : ACG_ERR_CUDA;
}

int acgb200_host_unregister(void *ptr)
{
    return cudaHostUnregister(ptr) == cudaSuccess ? ACG_SUCCESS : ACG_ERR_CUDA;
}

int acgb200_nccl_unique_id(void *id128)
{
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) return ACG_ERR_NCCL;
    memcpy(id128, &id, sizeof(id));
    return ACG_SUCCESS;
}

int acgb200_comm_init_rank(struct acgcomm *comm, int nranks, const void *id128, int rank, int *ncclerrcode)
{
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t nc;
    ncclResult_t r = ncclCommInitRank(&nc, nranks, id, rank);
    if (r != ncclSuccess) { if (ncclerrcode) *ncclerrcode = (int) r; return ACG_ERR_NCCL; }
    return acgcomm_init_nccl(comm, nc, ncclerrcode);
}

int acgb200_comm_destroy(struct acgcomm *comm)
{
    if (comm->type == acgcomm_nccl && comm->ncclcomm) ncclCommDestroy(comm->ncclcomm);
    acgcomm_free(comm);
    return ACG_SUCCESS;
}
