/*
 * mtxfile.c -- Matrix Market ingest for the solver's own matrix type, including
 * a per-rank read that never materialises the global matrix.
 *
 * The reference driver reads the whole file on rank 0 (acg/mtxfile.c, called
 * from cuda/acg-cuda.c:1297-1304), partitions there and scatters the parts
 * (:1516-1782); a 448^3 27-point matrix is 19 GB of file and as much again in
 * memory on that one rank.  Here every rank streams the file once and keeps
 * only the entries with an end among its own rows:
 *
 *   acgb200_mtx_info       header + size line (both encodings)
 *   acgb200_mtx_read       whole matrix -> acgsymcsrmatrix (text or binary)
 *   acgb200_mtx_read_part  one part of a row partition, binary encoding
 *
 * Encodings: the text format, and aCG's binary variant -- the text header and
 * size line followed by rowidx[nnz], colidx[nnz] (1-based acgidx_t) and
 * a[nnz] (double), acg/mtxfile.c:1107-1127, as written by mtx2bin
 * (mtx2bin/mtx2bin.c:538-549).  Only "matrix coordinate real symmetric" is
 * accepted: it is what the solver takes (acg/symcsrmatrix.c:66).
 */
#include "acgb200/error.h"
#include "acgb200/ext.h"
#include "acgb200/symcsrmatrix.h"
#include "hostmem.h"

#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define CHUNK ((int64_t) 1 << 22)      /* entries per read */

int acgb200_mtx_info(const char *path, struct acgb200_mtxinfo *info)
{
    memset(info, 0, sizeof(*info));
    FILE *f = fopen(path, "rb");
    if (!f) return ACG_ERR_ERRNO;
    char line[1100];
    int err = ACG_ERR_EOF;
    if (!fgets(line, sizeof(line), f)) goto done;
    {
        char object[64], format[64], field[64], symmetry[64];
        if (sscanf(line, "%%%%MatrixMarket %63s %63s %63s %63s", object, format, field, symmetry) != 4) { err = ACG_ERR_INVALID_VALUE; goto done; }
        for (char *p = field; *p; p++) *p = (char) tolower((unsigned char) *p);
        for (char *p = symmetry; *p; p++) *p = (char) tolower((unsigned char) *p);
        if (strcmp(object, "matrix") || strcmp(format, "coordinate")) { err = ACG_ERR_NOT_SUPPORTED; goto done; }
        info->field = !strcmp(field, "real") ? 0 : !strcmp(field, "integer") ? 1 : !strcmp(field, "pattern") ? 2 : 3;
        info->symmetric = !strcmp(symmetry, "symmetric");
    }
    for (;;) {
        if (!fgets(line, sizeof(line), f)) goto done;
        if (!strchr(line, '\n') && !feof(f)) { err = ACG_ERR_LINE_TOO_LONG; goto done; }
        if (line[0] != '%') break;
    }
    {
        long long nr, nc, nz;
        if (sscanf(line, "%lld %lld %lld", &nr, &nc, &nz) != 3 || nr < 0 || nc < 0 || nz < 0) { err = ACG_ERR_INVALID_VALUE; goto done; }
        info->nrows = nr; info->ncols = nc; info->nnzs = nz;
    }
    info->data_offset = (int64_t) ftello(f);
    err = ACG_SUCCESS;
done:
    fclose(f);
    return err;
}

static int check_info(const struct acgb200_mtxinfo *info)
{
    if (info->field != 0 || !info->symmetric) return ACG_ERR_NOT_SUPPORTED;
    if (info->nrows != info->ncols) return ACG_ERR_INVALID_VALUE;
    if (info->nrows > ACGIDX_T_MAX) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    return ACG_SUCCESS;
}

static int pread_all(int fd, void *buf, size_t bytes, int64_t off)
{
    char *p = buf;
    while (bytes > 0) {
        const ssize_t r = pread(fd, p, bytes, (off_t) off);
        if (r < 0) { if (errno == EINTR) continue; return ACG_ERR_ERRNO; }
        if (r == 0) return ACG_ERR_EOF;
        p += r; off += r; bytes -= (size_t) r;
    }
    return ACG_SUCCESS;
}

int acgb200_mtx_read(const char *path, int binary, struct acgsymcsrmatrix *A)
{
    struct acgb200_mtxinfo info;
    int err = acgb200_mtx_info(path, &info);
    if (err) return err;
    if ((err = check_info(&info))) return err;
    const int64_t nnz = info.nnzs;
    acgidx_t *ri = acgb200_bigalloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*ri));
    acgidx_t *ci = acgb200_bigalloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*ci));
    double *va = acgb200_bigalloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*va));
    if (!ri || !ci || !va) { err = ACG_ERR_ERRNO; goto done; }
    if (binary) {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { err = ACG_ERR_ERRNO; goto done; }
        err = pread_all(fd, ri, (size_t) nnz * sizeof(*ri), info.data_offset);
        if (!err) err = pread_all(fd, ci, (size_t) nnz * sizeof(*ci), info.data_offset + nnz * (int64_t) sizeof(*ri));
        if (!err) err = pread_all(fd, va, (size_t) nnz * sizeof(*va), info.data_offset + nnz * (int64_t) (sizeof(*ri) + sizeof(*ci)));
        close(fd);
        if (err) goto done;
    } else {
        FILE *f = fopen(path, "rb");
        if (!f) { err = ACG_ERR_ERRNO; goto done; }
        if (fseeko(f, (off_t) info.data_offset, SEEK_SET)) { fclose(f); err = ACG_ERR_ERRNO; goto done; }
        char line[1100];
        for (int64_t k = 0; k < nnz; k++) {
            if (!fgets(line, sizeof(line), f)) { err = ACG_ERR_EOF; break; }
            char *s = line, *t;
            const long long i = strtoll(s, &t, 10); if (t == s) { err = ACG_ERR_INVALID_VALUE; break; } s = t;
            const long long j = strtoll(s, &t, 10); if (t == s) { err = ACG_ERR_INVALID_VALUE; break; } s = t;
            const double x = strtod(s, &t); if (t == s) { err = ACG_ERR_INVALID_VALUE; break; }
            ri[k] = (acgidx_t) i; ci[k] = (acgidx_t) j; va[k] = x;
        }
        fclose(f);
        if (err) goto done;
    }
    for (int64_t k = 0; k < nnz; k++) {
        if (ri[k] < 1 || ri[k] > info.nrows || ci[k] < 1 || ci[k] > info.ncols) { err = ACG_ERR_INDEX_OUT_OF_BOUNDS; goto done; }
        ri[k]--; ci[k]--;                        /* the solver's matrices are 0-based */
    }
    err = acgsymcsrmatrix_init_real_double(A, (acgidx_t) info.nrows, nnz, 0, ri, ci, va);
done:
    free(ri); free(ci); free(va);
    return err;
}

struct kept { int64_t n; acgidx_t *ri, *ci; double *va; };

int acgb200_mtx_read_part(const char *path, int nparts, const int *rowparts, int part, struct acgsymcsrmatrix *A)
{
    struct acgb200_mtxinfo info;
    int err = acgb200_mtx_info(path, &info);
    if (err) return err;
    if ((err = check_info(&info))) return err;
    if (nparts < 1 || part < 0 || part >= nparts) return ACG_ERR_INVALID_VALUE;
    const int64_t nnz = info.nnzs, n = info.nrows;
    for (int64_t i = 0; i < n; i++)
        if (rowparts[i] < 0 || rowparts[i] >= nparts) return ACG_ERR_INDEX_OUT_OF_BOUNDS;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return ACG_ERR_ERRNO;
    const int64_t off_r = info.data_offset, off_c = off_r + nnz * (int64_t) sizeof(acgidx_t),
                  off_a = off_c + nnz * (int64_t) sizeof(acgidx_t);
    const int64_t nchunks = (nnz + CHUNK - 1) / CHUNK;
    struct kept *kp = calloc((size_t) (nchunks > 0 ? nchunks : 1), sizeof(*kp));
    if (!kp) { close(fd); return ACG_ERR_ERRNO; }
    int fail = 0;

    /* stream the three arrays chunk by chunk; keep the entries that touch an owned row */
#pragma omp parallel
    {
        acgidx_t *ri = malloc((size_t) CHUNK * sizeof(*ri));
        acgidx_t *ci = malloc((size_t) CHUNK * sizeof(*ci));
        double *va = malloc((size_t) CHUNK * sizeof(*va));
#pragma omp for schedule(dynamic)
        for (int64_t c = 0; c < nchunks; c++) {
            int bad = __atomic_load_n(&fail, __ATOMIC_RELAXED);
            if (bad || !ri || !ci || !va) { if (!bad) __atomic_store_n(&fail, ACG_ERR_ERRNO, __ATOMIC_RELAXED); continue; }
            const int64_t k0 = c * CHUNK, m = (nnz - k0 < CHUNK) ? nnz - k0 : CHUNK;
            int e = pread_all(fd, ri, (size_t) m * sizeof(*ri), off_r + k0 * (int64_t) sizeof(*ri));
            if (!e) e = pread_all(fd, ci, (size_t) m * sizeof(*ci), off_c + k0 * (int64_t) sizeof(*ci));
            int64_t keep = 0;
            for (int64_t k = 0; k < m && !e; k++) {
                if (ri[k] < 1 || ri[k] > n || ci[k] < 1 || ci[k] > n) { e = ACG_ERR_INDEX_OUT_OF_BOUNDS; break; }
                keep += rowparts[ri[k] - 1] == part || rowparts[ci[k] - 1] == part;
            }
            if (!e && keep > 0) {
                e = pread_all(fd, va, (size_t) m * sizeof(*va), off_a + k0 * (int64_t) sizeof(*va));
                struct kept *q = &kp[c];
                q->ri = malloc((size_t) keep * sizeof(*q->ri));
                q->ci = malloc((size_t) keep * sizeof(*q->ci));
                q->va = malloc((size_t) keep * sizeof(*q->va));
                if (!q->ri || !q->ci || !q->va) e = e ? e : ACG_ERR_ERRNO;
                for (int64_t k = 0; k < m && !e; k++) {
                    if (rowparts[ri[k] - 1] == part || rowparts[ci[k] - 1] == part) {
                        q->ri[q->n] = ri[k] - 1; q->ci[q->n] = ci[k] - 1; q->va[q->n] = va[k]; q->n++;
                    }
                }
            }
            if (e) __atomic_store_n(&fail, e, __ATOMIC_RELAXED);
        }
        free(ri); free(ci); free(va);
    }
    close(fd);
    err = fail;

    /* concatenate in file order (so the result does not depend on the thread count) */
    int64_t tot = 0;
    for (int64_t c = 0; c < nchunks; c++) tot += kp[c].n;
    acgidx_t *ri = NULL, *ci = NULL;
    double *va = NULL;
    if (!err) {
        ri = acgb200_bigalloc((size_t) (tot > 0 ? tot : 1) * sizeof(*ri));
        ci = acgb200_bigalloc((size_t) (tot > 0 ? tot : 1) * sizeof(*ci));
        va = acgb200_bigalloc((size_t) (tot > 0 ? tot : 1) * sizeof(*va));
        if (!ri || !ci || !va) err = ACG_ERR_ERRNO;
    }
    int64_t pos = 0;
    for (int64_t c = 0; c < nchunks; c++) {
        if (!err && kp[c].n > 0) {
            memcpy(ri + pos, kp[c].ri, (size_t) kp[c].n * sizeof(*ri));
            memcpy(ci + pos, kp[c].ci, (size_t) kp[c].n * sizeof(*ci));
            memcpy(va + pos, kp[c].va, (size_t) kp[c].n * sizeof(*va));
            pos += kp[c].n;
        }
        free(kp[c].ri); free(kp[c].ci); free(kp[c].va);
    }
    free(kp);

    /* the kept entries form a matrix on all n rows in which every row of `part`
     * is complete; partitioning it yields that part exactly (the other parts
     * come out truncated and are dropped) */
    struct acgsymcsrmatrix R, *sub = NULL;
    memset(&R, 0, sizeof(R));
    if (!err) err = acgsymcsrmatrix_init_real_double(&R, (acgidx_t) n, tot, 0, ri, ci, va);
    free(ri); free(ci); free(va);
    if (!err) {
        sub = calloc((size_t) nparts, sizeof(*sub));
        if (!sub) err = ACG_ERR_ERRNO;
    }
    if (!err) err = acgsymcsrmatrix_partition(&R, nparts, rowparts, sub, 0);
    if (!err) {
        for (int p = 0; p < nparts; p++) if (p != part) acgsymcsrmatrix_free(&sub[p]);
        *A = sub[part];
        /* counters that describe the whole matrix, not the truncated one */
        A->nnzs = nnz;
        if (A->graph) A->graph->nedges = nnz;
    }
    free(sub);
    acgsymcsrmatrix_free(&R);
    return err;
}
