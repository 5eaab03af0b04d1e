/*
 * mpi.h -- single-process stand-in for MPI (this image ships no MPI).
 *
 * Purpose: let the UNMODIFIED reference driver cuda/acg-cuda.c, which calls
 * MPI unconditionally (MPI_Init_thread at :891, bootstrap broadcast of the
 * NCCL id at :1113, matrix scatter, result gather), be compiled and linked
 * against libacgb200 for one-rank runs, and let libacgb200 be built with
 * -DACG_HAVE_MPI so that struct acgcomm has the layout the driver uses
 * (acg/comm.h:103-117).  Every communicator has size 1; collectives copy
 * in to out; point-to-point calls are errors.  Not a general MPI.
 *
 * On a machine with a real MPI this directory is simply left off the include
 * path.
 */
#ifndef ACGB200_MPI_SHIM_H
#define ACGB200_MPI_SHIM_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

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
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_INFO_NULL 0
#define MPI_COMM_TYPE_SHARED 1
#define MPI_IN_PLACE ((void *) 1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_REQUEST_NULL 0
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_LIBRARY_VERSION_STRING 256
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

/* datatypes: the value is the element size, derived types get 0x1000+size */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 0x101
#define MPI_C_BOOL 0x201
#define MPI_INT 4
#define MPI_INT32_T 0x104
#define MPI_INT64_T 8
#define MPI_DOUBLE 0x108
#define MPI_2INT 0x208
#define ACGB200_MPI_SIZEOF(t) ((t) >= 0x1000 ? (t) - 0x1000 : ((t) & 0xff))

#define MPI_OP_NULL 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_LOR 3
#define MPI_MAXLOC 4

static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided)
{ (void) argc; (void) argv; if (provided) *provided = required; return MPI_SUCCESS; }
static inline int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Query_thread(int *provided) { *provided = MPI_THREAD_FUNNELED; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm comm, int code) { (void) comm; exit(code ? code : 1); return MPI_SUCCESS; }
static inline double MPI_Wtime(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec; }
static inline int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) { *out = comm; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Comm_split_type(MPI_Comm comm, int type, int key, MPI_Info info, MPI_Comm *out)
{ (void) type; (void) key; (void) info; *out = comm; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm comm) { (void) comm; return MPI_SUCCESS; }
static inline int MPI_Get_processor_name(char *name, int *len)
{ strcpy(name, "localhost"); *len = 9; return MPI_SUCCESS; }
static inline int MPI_Get_library_version(char *version, int *len)
{ strcpy(version, "acgb200 single-process MPI shim"); *len = (int) strlen(version); return MPI_SUCCESS; }
static inline int MPI_Error_string(int err, char *s, int *len)
{ snprintf(s, MPI_MAX_ERROR_STRING, "MPI shim error %d", err); *len = (int) strlen(s); return MPI_SUCCESS; }

static inline int MPI_Type_size(MPI_Datatype t, int *size) { *size = ACGB200_MPI_SIZEOF(t); return MPI_SUCCESS; }
static inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newtype)
{ *newtype = 0x1000 + count * ACGB200_MPI_SIZEOF(old); return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void) t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *t) { *t = MPI_DATATYPE_NULL; return MPI_SUCCESS; }

static inline int acgb200_mpi_copy(const void *src, void *dst, int count, MPI_Datatype t)
{
    if (src != MPI_IN_PLACE && src != dst && count > 0)
        memmove(dst, src, (size_t) count * (size_t) ACGB200_MPI_SIZEOF(t));
    return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm)
{ (void) buf; (void) count; (void) t; (void) root; (void) comm; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{ (void) op; (void) comm; return acgb200_mpi_copy(s, r, count, t); }
static inline int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm)
{ (void) op; (void) root; (void) comm; return acgb200_mpi_copy(s, r, count, t); }
static inline int MPI_Exscan(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{ (void) s; (void) r; (void) count; (void) t; (void) op; (void) comm; return MPI_SUCCESS; }   /* rank 0's result is undefined */
static inline int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, int root, MPI_Comm comm)
{ (void) rc; (void) rt; (void) root; (void) comm; return acgb200_mpi_copy(s, r, sc, st); }
static inline int MPI_Gatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rcs, const int *displs,
                              MPI_Datatype rt, int root, MPI_Comm comm)
{ (void) rcs; (void) root; (void) comm;
  return acgb200_mpi_copy(s, (char *) r + (size_t) (displs ? displs[0] : 0) * ACGB200_MPI_SIZEOF(rt), sc, st); }
static inline int MPI_Scatterv(const void *s, const int *scs, const int *displs, MPI_Datatype st, void *r, int rc,
                               MPI_Datatype rt, int root, MPI_Comm comm)
{ (void) scs; (void) root; (void) comm;
  return acgb200_mpi_copy((const char *) s + (size_t) (displs ? displs[0] : 0) * ACGB200_MPI_SIZEOF(st), r, rc, rt); }

/* point-to-point: there is no other rank */
static inline int MPI_Send(const void *b, int c, MPI_Datatype t, int d, int tag, MPI_Comm comm)
{ (void) b; (void) c; (void) t; (void) d; (void) tag; (void) comm; return MPI_ERR_OTHER; }
static inline int MPI_Recv(void *b, int c, MPI_Datatype t, int s, int tag, MPI_Comm comm, MPI_Status *st)
{ (void) b; (void) c; (void) t; (void) s; (void) tag; (void) comm; (void) st; return MPI_ERR_OTHER; }
static inline int MPI_Isend(const void *b, int c, MPI_Datatype t, int d, int tag, MPI_Comm comm, MPI_Request *r)
{ (void) b; (void) c; (void) t; (void) d; (void) tag; (void) comm; *r = MPI_REQUEST_NULL; return MPI_ERR_OTHER; }
static inline int MPI_Irecv(void *b, int c, MPI_Datatype t, int s, int tag, MPI_Comm comm, MPI_Request *r)
{ (void) b; (void) c; (void) t; (void) s; (void) tag; (void) comm; *r = MPI_REQUEST_NULL; return MPI_ERR_OTHER; }
static inline int MPI_Send_init(const void *b, int c, MPI_Datatype t, int d, int tag, MPI_Comm comm, MPI_Request *r)
{ (void) b; (void) c; (void) t; (void) d; (void) tag; (void) comm; *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Recv_init(void *b, int c, MPI_Datatype t, int s, int tag, MPI_Comm comm, MPI_Request *r)
{ (void) b; (void) c; (void) t; (void) s; (void) tag; (void) comm; *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Startall(int n, MPI_Request *r) { (void) n; (void) r; return MPI_SUCCESS; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void) r; (void) s; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void) n; (void) r; (void) s; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request *r, int *flag, MPI_Status *s) { (void) r; (void) s; *flag = 1; return MPI_SUCCESS; }

#ifdef __cplusplus
}
#endif
#endif
