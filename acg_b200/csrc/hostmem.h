/*
 * hostmem.h -- allocation of the large host arrays (full-storage CSR, packed
 * edges): free()-compatible (the structs are the reference's, and whoever owns
 * them may release the arrays with plain free()), 2 MiB-aligned and advised
 * for transparent huge pages, so that the first touch of a multi-gigabyte
 * array costs thousands of page faults instead of millions -- on the 27-point
 * 224^3 matrix the first-touch faults were most of acgsymcsrmatrix_dsymv_init.
 */
#ifndef ACGB200_HOSTMEM_H
#define ACGB200_HOSTMEM_H

#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

static inline void *acgb200_bigalloc(size_t bytes)
{
    if (bytes < ((size_t) 8 << 20)) return malloc(bytes > 0 ? bytes : 1);
    void *p = NULL;
    if (posix_memalign(&p, (size_t) 2 << 20, bytes)) return NULL;
#ifdef MADV_HUGEPAGE
    madvise(p, bytes, MADV_HUGEPAGE);      /* advisory: failure is harmless */
#endif
    return p;
}

static inline void *acgb200_bigcalloc(size_t count, size_t size)
{
    if (size && count > (size_t) -1 / size) return NULL;
    void *p = acgb200_bigalloc(count * size);
    if (p) memset(p, 0, count * size);
    return p;
}

#endif
