/*
 * acgb200/error.h -- error codes; values identical to the reference's
 * enum acgerrcode (acg/error.h:47-104) because they are the return
 * convention of every entry point and the driver prints them with
 * acgerrcodestr() (cuda/acg-cuda.c, every call site).
 */
#ifndef ACGB200_ERROR_H
#define ACGB200_ERROR_H

#ifdef __cplusplus
extern "C" {
#endif

enum acgerrcode {
    ACG_SUCCESS = 0,
    ACG_ERR_ERRNO, ACG_ERR_FEXCEPT, ACG_ERR_MPI, ACG_ERR_CUDA, ACG_ERR_NCCL,
    ACG_ERR_NVSHMEM, ACG_ERR_CUBLAS, ACG_ERR_CUSPARSE, ACG_ERR_HIP, ACG_ERR_RCCL,
    ACG_ERR_ROCSHMEM, ACG_ERR_HIPBLAS, ACG_ERR_HIPSPARSE,
    ACG_ERR_MPI_NOT_SUPPORTED, ACG_ERR_NCCL_NOT_SUPPORTED,
    ACG_ERR_NVSHMEM_NOT_SUPPORTED, ACG_ERR_RCCL_NOT_SUPPORTED,
    ACG_ERR_ROCSHMEM_NOT_SUPPORTED, ACG_ERR_METIS_NOT_SUPPORTED,
    ACG_ERR_PETSC_NOT_SUPPORTED, ACG_ERR_LIBZ_NOT_SUPPORTED,
    ACG_ERR_METIS_INPUT, ACG_ERR_METIS_MEMORY, ACG_ERR_METIS, ACG_ERR_METIS_EOVERFLOW,
    ACG_ERR_NOT_SUPPORTED, ACG_ERR_EOF, ACG_ERR_LINE_TOO_LONG, ACG_ERR_INVALID_VALUE,
    ACG_ERR_OVERFLOW, ACG_ERR_INDEX_OUT_OF_BOUNDS, ACG_ERR_NO_BUFFER_SPACE,
    ACG_ERR_MTX_INVALID_COMMENT, ACG_ERR_INVALID_FORMAT_SPECIFIER,
    ACG_ERR_VECTOR_INCOMPATIBLE_SIZE, ACG_ERR_VECTOR_INCOMPATIBLE_FORMAT,
    ACG_ERR_VECTOR_EXPECTED_FULL, ACG_ERR_VECTOR_EXPECTED_PACKED,
    ACG_ERR_NOT_CONVERGED, ACG_ERR_NOT_CONVERGED_INDEFINITE_MATRIX,
};

/* replaces acgerrcodestr (acg/error.h:113, acg/error.c); the second argument
 * carries the third-party code (CUDA/NCCL/MPI) when err names one */
const char *acgerrcodestr(int err, int mpierrcode);

#ifdef __cplusplus
}
#endif
#endif
