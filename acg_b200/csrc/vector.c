/*
 * vector.c -- host-side vectors with a trailing ghost segment.
 *
 * Own implementation of the acgvector_* subset declared in
 * include/acgb200/vector.h (reference: acg/vector.c).  These are setup /
 * diagnostic helpers; the solver's BLAS-1 runs on the device (kernels.cu).
 */
#include "acgb200/error.h"
#include "acgb200/vector.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static acgidx_t owned(const struct acgvector *x)
{
    /* acg/vector.c:468: packed vectors exclude the ghost tail, full ones have none */
    return x->idx ? x->num_nonzeros - x->num_ghost_nonzeros : x->size;
}

void acgvector_init_empty(struct acgvector *x)
{
    memset(x, 0, sizeof(*x));
    x->nparts = x->nprocs = x->npparts = 1;
}

void acgvector_free(struct acgvector *x)
{
    free(x->x); free(x->idx);
    x->x = NULL; x->idx = NULL;
}

int acgvector_alloc(struct acgvector *x, acgidx_t size)
{
    acgvector_init_empty(x);
    x->size = size;
    x->num_nonzeros = size;
    x->x = malloc((size_t) (size > 0 ? size : 1) * sizeof(*x->x));
    return x->x ? ACG_SUCCESS : ACG_ERR_ERRNO;
}

int acgvector_init_real_double(struct acgvector *x, acgidx_t size, const double *data)
{
    int err = acgvector_alloc(x, size);
    if (err) return err;
    memcpy(x->x, data, (size_t) size * sizeof(*x->x));
    return ACG_SUCCESS;
}

int acgvector_alloc_packed(struct acgvector *x, acgidx_t size, acgidx_t num_nonzeros, int idxbase, const acgidx_t *idx)
{
    acgvector_init_empty(x);
    x->size = size;
    x->num_nonzeros = num_nonzeros;
    x->idxbase = idxbase;
    size_t m = (size_t) (num_nonzeros > 0 ? num_nonzeros : 1);
    x->x = malloc(m * sizeof(*x->x));
    x->idx = malloc(m * sizeof(*x->idx));
    if (!x->x || !x->idx) { acgvector_free(x); return ACG_ERR_ERRNO; }
    if (idx) memcpy(x->idx, idx, (size_t) num_nonzeros * sizeof(*x->idx));
    return ACG_SUCCESS;
}

int acgvector_init_copy(struct acgvector *dst, const struct acgvector *src)
{
    *dst = *src;
    size_t m = (size_t) (src->num_nonzeros > 0 ? src->num_nonzeros : 1);
    dst->x = malloc(m * sizeof(*dst->x));
    if (!dst->x) return ACG_ERR_ERRNO;
    memcpy(dst->x, src->x, (size_t) src->num_nonzeros * sizeof(*dst->x));
    dst->idx = NULL;
    if (src->idx) {
        dst->idx = malloc(m * sizeof(*dst->idx));
        if (!dst->idx) { free(dst->x); return ACG_ERR_ERRNO; }
        memcpy(dst->idx, src->idx, (size_t) src->num_nonzeros * sizeof(*dst->idx));
    }
    return ACG_SUCCESS;
}

int acgvector_setzero(struct acgvector *x)
{
    memset(x->x, 0, (size_t) x->num_nonzeros * sizeof(*x->x));
    return ACG_SUCCESS;
}

int acgvector_set_constant_real_double(struct acgvector *x, double a)
{
    for (acgidx_t k = 0; k < x->num_nonzeros; k++) x->x[k] = a;
    return ACG_SUCCESS;
}

int acgvector_copy(struct acgvector *y, const struct acgvector *x, int64_t *num_bytes)
{
    if (x->size != y->size || owned(x) != owned(y)) return ACG_ERR_VECTOR_INCOMPATIBLE_SIZE;
    memcpy(y->x, x->x, (size_t) owned(x) * sizeof(*y->x));
    if (num_bytes) *num_bytes += (int64_t) owned(x) * 2 * (int64_t) sizeof(double);
    return ACG_SUCCESS;
}

int acgvector_daxpy(double a, const struct acgvector *x, struct acgvector *y, int64_t *num_flops, int64_t *num_bytes)
{
    if (x->size != y->size || owned(x) != owned(y)) return ACG_ERR_VECTOR_INCOMPATIBLE_SIZE;
    const acgidx_t n = owned(x);
    for (acgidx_t k = 0; k < n; k++) y->x[k] += a * x->x[k];
    if (num_flops) *num_flops += 2 * (int64_t) n;
    if (num_bytes) *num_bytes += (int64_t) n * 2 * (int64_t) sizeof(double);
    return ACG_SUCCESS;
}

int acgvector_dnrm2(const struct acgvector *x, double *nrm2, int64_t *num_flops, int64_t *num_bytes)
{
    const acgidx_t n = owned(x);
    double c = 0;
    for (acgidx_t k = 0; k < n; k++) c += x->x[k] * x->x[k];
    *nrm2 = sqrt(c);
    if (num_flops) *num_flops += 2 * (int64_t) n + 1;
    if (num_bytes) *num_bytes += (int64_t) n * (int64_t) sizeof(double);
    return ACG_SUCCESS;
}

int acgvector_usga(struct acgvector *x, const struct acgvector *y)
{
    /* acg/vector.c:767: gather a full vector into a packed one */
    if (!x->idx) return ACG_ERR_VECTOR_EXPECTED_PACKED;
    if (y->idx) return ACG_ERR_VECTOR_EXPECTED_FULL;
    if (x->size != y->size) return ACG_ERR_VECTOR_INCOMPATIBLE_SIZE;
    for (acgidx_t k = 0; k < x->num_nonzeros; k++) x->x[k] = y->x[x->idx[k] - x->idxbase];
    return ACG_SUCCESS;
}

int acgvector_ussc(struct acgvector *y, const struct acgvector *x)
{
    /* acg/vector.c:811: scatter a packed vector into a full one */
    if (!x->idx) return ACG_ERR_VECTOR_EXPECTED_PACKED;
    if (y->idx) return ACG_ERR_VECTOR_EXPECTED_FULL;
    if (x->size != y->size) return ACG_ERR_VECTOR_INCOMPATIBLE_SIZE;
    for (acgidx_t k = 0; k < x->num_nonzeros; k++) y->x[x->idx[k] - x->idxbase] = x->x[k];
    return ACG_SUCCESS;
}
