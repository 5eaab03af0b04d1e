/*
 * driver_mock.c -- TEST INFRASTRUCTURE ONLY (see cuda_mock.c).  The few extra CUDA-runtime,
 * cuBLAS, cuSPARSE and NCCL entry points the UNMODIFIED reference driver cuda/acg-cuda.c calls
 * around the solver (device selection, handle creation, version strings), so that the driver
 * objects built by tools/build_driver.sh can be linked against the device stand-in and the whole
 * drop-in boundary -- driver -> acgsolvercuda_init / _solvempi / _solve_pipelined / _fwritempi ->
 * report and solution -- can run in the CPU test-suite.
 */
#include <cublas_v2.h>
#include <cuda_runtime_api.h>
#include <cusparse.h>
#include <nccl.h>

cudaError_t cudaSetDevice(int d) { (void) d; return cudaSuccess; }
cudaError_t cudaDeviceReset(void) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaRuntimeGetVersion(int *v) { *v = 12090; return cudaSuccess; }

cublasStatus_t cublasCreate_v2(cublasHandle_t *h) { *h = (cublasHandle_t) 0x1; return CUBLAS_STATUS_SUCCESS; }
cublasStatus_t cublasDestroy_v2(cublasHandle_t h) { (void) h; return CUBLAS_STATUS_SUCCESS; }
const char *cublasGetStatusString(cublasStatus_t s) { (void) s; return "host-simulation stand-in"; }

cusparseStatus_t cusparseCreate(cusparseHandle_t *h) { *h = (cusparseHandle_t) 0x1; return CUSPARSE_STATUS_SUCCESS; }
cusparseStatus_t cusparseDestroy(cusparseHandle_t h) { (void) h; return CUSPARSE_STATUS_SUCCESS; }
const char *cusparseGetErrorString(cusparseStatus_t s) { (void) s; return "host-simulation stand-in"; }
cusparseStatus_t cusparseGetVersion(cusparseHandle_t h, int *v) { (void) h; *v = 12000; return CUSPARSE_STATUS_SUCCESS; }

ncclResult_t ncclGetVersion(int *v) { *v = 22809; return ncclSuccess; }
