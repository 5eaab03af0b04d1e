/*
 * acgb200/vector.h -- dense / packed vectors with a trailing ghost segment.
 *
 * ABI counterpart of acg/vector.h:58-161 (struct acgvector) and of the
 * subset of acgvector_* entry points that the solver path and the driver
 * cuda/acg-cuda.c touch.  Field order and types are the reference's: the
 * driver stack-allocates these structs and passes them by pointer.
 *
 * Layout contract used by the device path: x[0 .. num_nonzeros -
 * num_ghost_nonzeros) are owned entries, the last num_ghost_nonzeros are
 * ghosts (acg/vector.h:150-160); idx == NULL means full storage.
 */
#ifndef ACGB200_VECTOR_H
#define ACGB200_VECTOR_H

#include "acgb200/config.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct acgvector {
    /* partition bookkeeping (acg/vector.h:64-103) */
    int nparts, parttag, nprocs, npparts, ownerrank, ownerpart;
    /* elements (acg/vector.h:108-120) */
    acgidx_t size;
    double *x;
    /* packed storage (acg/vector.h:125-147) */
    acgidx_t num_nonzeros;
    int idxbase;
    acgidx_t *idx;
    /* ghosts are stored last (acg/vector.h:150-160) */
    acgidx_t num_ghost_nonzeros;
};

/* acg/vector.h:170 */ ACG_API void acgvector_init_empty(struct acgvector *x);
/* acg/vector.h:176 */ ACG_API void acgvector_free(struct acgvector *x);
/* acg/vector.h:191 */ ACG_API int acgvector_init_copy(struct acgvector *dst, const struct acgvector *src);
/* acg/vector.h:202 */ ACG_API int acgvector_alloc(struct acgvector *x, acgidx_t size);
/* acg/vector.h:210 */ ACG_API int acgvector_init_real_double(struct acgvector *x, acgidx_t size, const double *data);
/* acg/vector.h:223 */ ACG_API int acgvector_alloc_packed(struct acgvector *x, acgidx_t size, acgidx_t num_nonzeros, int idxbase, const acgidx_t *idx);
/* acg/vector.h:249 */ ACG_API int acgvector_setzero(struct acgvector *x);
/* acg/vector.h:256 */ ACG_API int acgvector_set_constant_real_double(struct acgvector *x, double a);

/* host BLAS-1 on the owned prefix -- used by the driver for the
 * manufactured-solution error norms (cuda/acg-cuda.c:2377-2385), never by the
 * solver loop, which runs on the device */
/* acg/vector.h:326 */ ACG_API int acgvector_copy(struct acgvector *y, const struct acgvector *x, int64_t *num_bytes);
/* acg/vector.h:347 */ ACG_API int acgvector_daxpy(double a, const struct acgvector *x, struct acgvector *y, int64_t *num_flops, int64_t *num_bytes);
/* acg/vector.h:385 */ ACG_API int acgvector_dnrm2(const struct acgvector *x, double *nrm2, int64_t *num_flops, int64_t *num_bytes);

/* gather from a full vector into a packed one: x[k] = y[idx[k]] (acg/vector.h:460) */
ACG_API int acgvector_usga(struct acgvector *x, const struct acgvector *y);
/* scatter a packed vector into a full one: y[idx[k]] = x[k] (acg/vector.h:479) */
ACG_API int acgvector_ussc(struct acgvector *y, const struct acgvector *x);

#ifdef __cplusplus
}
#endif
#endif
