/*
 * mpi.h -- the subset of MPI the UNMODIFIED reference driver cuda/acg-cuda.c and the
 * reference's host layer (acg/{vector,symcsrmatrix,graph,halo,mtxfile,error}.c) use,
 * for an image that ships no MPI.  Implemented by compat/mpi/mpishim.c
 * (libacgb200mpishim.so): one process per rank on ONE node, point-to-point over
 * Unix-domain sockets, collectives on top of it (linear; they carry the bootstrap --
 * NCCL id broadcast, matrix scatter, result gather, report reduction -- never the
 * CG loop, which runs over NCCL / NVLink).  Ranks are started by
 * compat/mpi/acgb200-mpirun (or any launcher that exports RANK / WORLD_SIZE /
 * MASTER_PORT, e.g. `torchrun --no-python`); without such an environment
 * MPI_Init yields a single rank, so the same binary also runs alone.
 *
 * This is bootstrap plumbing for the drop-in boundary, not a
 * general MPI: intra-node only, no derived datatypes beyond contiguous, no
 * one-sided or file operations.  On a machine with a real MPI this directory is
 * simply left off the include path.
 */
#ifndef ACGB200_MPI_SHIM_H
#define ACGB200_MPI_SHIM_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION 3
#define MPI_SUBVERSION 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; long long nbytes_; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_ERR_TRUNCATE 14
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_INFO_NULL 0
#define MPI_COMM_TYPE_SHARED 1
#define MPI_IN_PLACE ((void *) 1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_REQUEST_NULL 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_LIBRARY_VERSION_STRING 256
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

/* datatypes: the low byte is the element size, the next one tells kinds of equal size apart;
 * contiguous derived types are 0x1000 + size in bytes */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 0x001
#define MPI_BYTE 0x101
#define MPI_C_BOOL 0x201
#define MPI_INT 0x004
#define MPI_INT32_T 0x104
#define MPI_UNSIGNED 0x204
#define MPI_FLOAT 0x304
#define MPI_INT64_T 0x008
#define MPI_DOUBLE 0x108
#define MPI_2INT 0x208
#define MPI_LONG 0x308
#define MPI_UINT64_T 0x408
#define ACGB200_MPI_SIZEOF(t) ((t) >= 0x1000 ? (t) - 0x1000 : ((t) & 0xff))

#define MPI_OP_NULL 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_LOR 3
#define MPI_MAXLOC 4
#define MPI_MIN 5
#define MPI_LAND 6

int MPI_Init(int *argc, char ***argv);
int MPI_Init_thread(int *argc, char ***argv, int required, int *provided);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Query_thread(int *provided);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
int MPI_Get_processor_name(char *name, int *len);
int MPI_Get_library_version(char *version, int *len);
int MPI_Error_string(int err, char *s, int *len);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_split_type(MPI_Comm comm, int type, int key, MPI_Info info, MPI_Comm *out);

int MPI_Type_size(MPI_Datatype t, int *size);
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Send_init(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv_init(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Start(MPI_Request *req);
int MPI_Startall(int n, MPI_Request *reqs);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status);
int MPI_Request_free(MPI_Request *req);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype t, int *count);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Exscan(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt,
               int root, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, const int *recvcounts,
                const int *displs, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype st, void *recvbuf,
                 int recvcount, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype st, void *recvbuf, int recvcount, MPI_Datatype rt,
                  MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
