/*
 * error.c -- error-code strings.  Replaces acgerrcodestr (acg/error.c); the
 * message texts are this library's own, the code values are the reference's.
 */
#include "acgb200/error.h"

#include <cuda_runtime_api.h>
#include <errno.h>
#include <nccl.h>
#include <string.h>

const char *acgerrcodestr(int err, int code)
{
    switch (err) {
    case ACG_SUCCESS: return "success";
    case ACG_ERR_ERRNO: return strerror(errno);
    case ACG_ERR_FEXCEPT: return "floating-point exception";
    case ACG_ERR_MPI: return "MPI error";
    case ACG_ERR_CUDA: return cudaGetErrorString((cudaError_t) (code ? code : (int) cudaPeekAtLastError()));
    case ACG_ERR_NCCL: return ncclGetErrorString((ncclResult_t) code);
    case ACG_ERR_NVSHMEM: return "NVSHMEM error";
    case ACG_ERR_CUBLAS: return "cuBLAS error";
    case ACG_ERR_CUSPARSE: return "cuSPARSE error";
    case ACG_ERR_HIP: return "HIP error";
    case ACG_ERR_RCCL: return "RCCL error";
    case ACG_ERR_ROCSHMEM: return "rocSHMEM error";
    case ACG_ERR_HIPBLAS: return "hipBLAS error";
    case ACG_ERR_HIPSPARSE: return "hipSPARSE error";
    case ACG_ERR_MPI_NOT_SUPPORTED: return "MPI not supported in this build (NCCL is the only data path)";
    case ACG_ERR_NCCL_NOT_SUPPORTED: return "NCCL not supported";
    case ACG_ERR_NVSHMEM_NOT_SUPPORTED: return "NVSHMEM not supported in this build (NCCL is the only data path)";
    case ACG_ERR_RCCL_NOT_SUPPORTED: return "RCCL not supported";
    case ACG_ERR_ROCSHMEM_NOT_SUPPORTED: return "rocSHMEM not supported";
    case ACG_ERR_METIS_NOT_SUPPORTED: return "METIS not supported";
    case ACG_ERR_PETSC_NOT_SUPPORTED: return "PETSc not supported";
    case ACG_ERR_LIBZ_NOT_SUPPORTED: return "zlib not supported";
    case ACG_ERR_METIS_INPUT: return "METIS: input error";
    case ACG_ERR_METIS_MEMORY: return "METIS: out of memory";
    case ACG_ERR_METIS: return "METIS: error";
    case ACG_ERR_METIS_EOVERFLOW: return "METIS: value too large for data type";
    case ACG_ERR_NOT_SUPPORTED: return "not supported";
    case ACG_ERR_EOF: return "unexpected end of file";
    case ACG_ERR_LINE_TOO_LONG: return "line too long";
    case ACG_ERR_INVALID_VALUE: return "invalid value";
    case ACG_ERR_OVERFLOW: return "value too large for data type";
    case ACG_ERR_INDEX_OUT_OF_BOUNDS: return "index out of bounds";
    case ACG_ERR_NO_BUFFER_SPACE: return "not enough buffer space";
    case ACG_ERR_MTX_INVALID_COMMENT: return "invalid Matrix Market comment line";
    case ACG_ERR_INVALID_FORMAT_SPECIFIER: return "invalid format specifier";
    case ACG_ERR_VECTOR_INCOMPATIBLE_SIZE: return "incompatible vector size";
    case ACG_ERR_VECTOR_INCOMPATIBLE_FORMAT: return "incompatible vector format";
    case ACG_ERR_VECTOR_EXPECTED_FULL: return "expected a vector in full storage";
    case ACG_ERR_VECTOR_EXPECTED_PACKED: return "expected a vector in packed storage";
    case ACG_ERR_NOT_CONVERGED: return "not converged";
    case ACG_ERR_NOT_CONVERGED_INDEFINITE_MATRIX: return "not converged (indefinite matrix)";
    default: return "unknown error";
    }
}
