/*
 * ref_shim.c -- flat C entry points over the UNMODIFIED aCG reference.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together
 * with the reference's own sources (taken where they lie under
 * /root/reference, never copied into this repository) into
 * oracle/_ref/libacgref.so by oracle/Makefile.  It exists so that Python
 * (ctypes) can drive the reference's CPU path -- acgsymcsrmatrix_* /
 * acgvector_* / acgsolver_* -- without replicating its struct layouts.
 *
 * Uses: pinning oracle/cg_oracle.c bit-for-bit, generating tests/golden/,
 * and bench.py's "reference" CPU baseline.
 */
#include "acg/config.h"
#include "acg/cg.h"
#include "acg/error.h"
#include "acg/graph.h"
#include "acg/halo.h"
#include "acg/symcsrmatrix.h"
#include "acg/vector.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ref_sizeof_idx(void) { return (int) sizeof(acgidx_t); }

static int build(struct acgsymcsrmatrix *A, int n, int64_t nnz,
                 const int *row, const int *col, const double *a, double eps)
{
    int err = acgsymcsrmatrix_init_real_double(A, n, nnz, 0, row, col, a);
    if (err) return err;
    err = acgsymcsrmatrix_dsymv_init(A, eps);
    if (err) { acgsymcsrmatrix_free(A); return err; }
    return 0;
}

/* full-storage CSR exactly as the reference builds it */
int ref_full_csr(int n, int64_t nnz, const int *row, const int *col,
                 const double *a, double eps,
                 int64_t *frowptr, int *fcolidx, double *fa, int64_t *fnnz)
{
    struct acgsymcsrmatrix A;
    int err = build(&A, n, nnz, row, col, a, eps);
    if (err) return err;
    memcpy(frowptr, A.frowptr, ((size_t) n + 1) * sizeof(*frowptr));
    memcpy(fcolidx, A.fcolidx, (size_t) A.fnpnzs * sizeof(*fcolidx));
    memcpy(fa, A.fa, (size_t) A.fnpnzs * sizeof(*fa));
    *fnnz = A.fnpnzs;
    acgsymcsrmatrix_free(&A);
    return 0;
}

static int wrapvec(const struct acgsymcsrmatrix *A, struct acgvector *v, const double *src)
{
    int err = acgsymcsrmatrix_vector(A, v);
    if (err) return err;
    if (src) memcpy(v->x, src, (size_t) A->nrows * sizeof(double));
    else acgvector_setzero(v);
    return 0;
}

/* y = alpha*A*x + beta*y through acgsymcsrmatrix_dsymv on full storage */
int ref_dsymv(int n, int64_t nnz, const int *row, const int *col, const double *a,
              double alpha, const double *x, double beta, double *y)
{
    struct acgsymcsrmatrix A;
    int err = build(&A, n, nnz, row, col, a, 0.0);
    if (err) return err;
    struct acgvector vx, vy;
    err = wrapvec(&A, &vx, x); if (err) return err;
    err = wrapvec(&A, &vy, y); if (err) return err;
    err = acgsymcsrmatrix_dsymv(alpha, &A, &vx, beta, &vy, NULL, NULL);
    memcpy(y, vy.x, (size_t) n * sizeof(double));
    acgvector_free(&vx); acgvector_free(&vy);
    acgsymcsrmatrix_free(&A);
    return err;
}

static struct acgvector borrow(int n, double *x)
{
    struct acgvector v;
    memset(&v, 0, sizeof(v));
    v.nparts = 1; v.nprocs = 1; v.npparts = 1;
    v.size = n; v.x = x; v.num_nonzeros = n; v.idx = NULL;
    return v;
}

double ref_ddot(int n, const double *x, const double *y)
{
    struct acgvector vx = borrow(n, (double *) x), vy = borrow(n, (double *) y);
    double d = 0; acgvector_ddot(&vx, &vy, &d, NULL, NULL); return d;
}
double ref_dnrm2sqr(int n, const double *x)
{
    struct acgvector vx = borrow(n, (double *) x);
    double d = 0; acgvector_dnrm2sqr(&vx, &d, NULL, NULL); return d;
}
void ref_daxpy(int n, double a, const double *x, double *y)
{
    struct acgvector vx = borrow(n, (double *) x), vy = borrow(n, y);
    acgvector_daxpy(a, &vx, &vy, NULL, NULL);
}
void ref_daypx(int n, double a, double *y, const double *x)
{
    struct acgvector vx = borrow(n, (double *) x), vy = borrow(n, y);
    acgvector_daypx(a, &vy, &vx, NULL, NULL);
}

/*
 * The reference CPU solver, acgsolver_solve (acg/cg.c:198), on the whole
 * matrix.  out[0..5] = niterations, bnrm2, r0nrm2, rnrm2, tsolve, tgemv.
 * Returns the acgerrcode of the solve.
 */
int ref_cg(int n, int64_t nnz, const int *row, const int *col, const double *a,
           double eps, const double *b, double *x, int maxits,
           double residualatol, double residualrtol, double *out)
{
    struct acgsymcsrmatrix A;
    int err = build(&A, n, nnz, row, col, a, eps);
    if (err) return err;
    struct acgvector vb, vx;
    err = wrapvec(&A, &vb, b); if (err) return err;
    err = wrapvec(&A, &vx, x); if (err) return err;
    struct acgsolver cg;
    err = acgsolver_init(&cg, &A);
    if (err) return err;
    err = acgsolver_solve(&cg, &A, &vb, &vx, maxits, 0.0, 0.0, residualatol, residualrtol);
    memcpy(x, vx.x, (size_t) n * sizeof(double));
    out[0] = cg.niterations; out[1] = cg.bnrm2; out[2] = cg.r0nrm2;
    out[3] = cg.rnrm2; out[4] = cg.tsolve; out[5] = cg.tgemv;
    acgsolver_free(&cg);
    acgvector_free(&vb); acgvector_free(&vx);
    acgsymcsrmatrix_free(&A);
    return err;
}

/*
 * Persistent variant for benchmarking: set the problem up once (the
 * reference's COO -> packed -> full-storage path, acg/symcsrmatrix.c:66,:760),
 * then time acgsolver_solve alone as often as needed.
 */
struct ref_problem {
    struct acgsymcsrmatrix A;
    struct acgsolver cg;
    struct acgvector b, x;
};

void *ref_setup(int n, int64_t nnz, const int *row, const int *col, const double *a, double eps)
{
    struct ref_problem *P = calloc(1, sizeof(*P));
    if (!P) return NULL;
    if (build(&P->A, n, nnz, row, col, a, eps)) { free(P); return NULL; }
    if (wrapvec(&P->A, &P->b, NULL) || wrapvec(&P->A, &P->x, NULL)) return NULL;
    if (acgsolver_init(&P->cg, &P->A)) return NULL;
    return P;
}


/*
 * NUMA placement of the full-storage arrays: acgsymcsrmatrix_dsymv_init fills them from one
 * thread, so on a multi-socket host every page of the matrix lands on that thread's node and
 * the OpenMP dsymv (acg/symcsrmatrix.c:902, static schedule over rows) streams most of it over
 * the socket interconnect.  Here each thread of a loop with the same static schedule
 * first-touches (and copies) the slice of frowptr / fcolidx / fa it will stream; the old
 * arrays are freed and the pointers swapped.  Contents are bit-identical -- placement only.
 * Returns 1 if the arrays were moved.
 */
int ref_place(void *handle)
{
    struct ref_problem *P = handle;
    struct acgsymcsrmatrix *A = &P->A;
    if (!A->frowptr || !A->fcolidx || !A->fa) return 0;
    const acgidx_t m = A->nprows - A->nghostrows;
    const int64_t nnz = A->fnpnzs;
    int64_t *rp = malloc(((size_t) A->nprows + 1) * sizeof(*rp));
    acgidx_t *ci = malloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*ci));
    double *va = malloc((size_t) (nnz > 0 ? nnz : 1) * sizeof(*va));
    if (!rp || !ci || !va) { free(rp); free(ci); free(va); return 0; }
    const int64_t *orp = A->frowptr;
    #pragma omp parallel for
    for (acgidx_t i = 0; i < m - m % 4; i += 4) {
        for (int r = 0; r < 4; r++) rp[i + r] = orp[i + r];
        for (int64_t k = orp[i]; k < orp[i + 4]; k++) { ci[k] = A->fcolidx[k]; va[k] = A->fa[k]; }
    }
    for (acgidx_t i = m - m % 4; i <= A->nprows; i++) rp[i] = orp[i];
    for (int64_t k = orp[m - m % 4]; k < nnz; k++) { ci[k] = A->fcolidx[k]; va[k] = A->fa[k]; }
    free(A->frowptr); free(A->fcolidx); free(A->fa);
    A->frowptr = rp; A->fcolidx = ci; A->fa = va;
    return 1;
}

/* out[0..5] as in ref_cg; x0 may be NULL (zero initial guess) */
int ref_solve(void *handle, const double *b, const double *x0, double *xout, int maxits,
              double residualatol, double residualrtol, double *out)
{
    struct ref_problem *P = handle;
    const size_t n = (size_t) P->A.nrows;
    memcpy(P->b.x, b, n * sizeof(double));
    if (x0) memcpy(P->x.x, x0, n * sizeof(double)); else memset(P->x.x, 0, n * sizeof(double));
    const double tsolve0 = P->cg.tsolve, tgemv0 = P->cg.tgemv;
    int err = acgsolver_solve(&P->cg, &P->A, &P->b, &P->x, maxits, 0.0, 0.0, residualatol, residualrtol);
    if (xout) memcpy(xout, P->x.x, n * sizeof(double));
    out[0] = P->cg.niterations; out[1] = P->cg.bnrm2; out[2] = P->cg.r0nrm2;
    out[3] = P->cg.rnrm2; out[4] = P->cg.tsolve - tsolve0; out[5] = P->cg.tgemv - tgemv0;
    return err;
}

void ref_free(void *handle)
{
    struct ref_problem *P = handle;
    if (!P) return;
    acgsolver_free(&P->cg);
    acgvector_free(&P->b); acgvector_free(&P->x);
    acgsymcsrmatrix_free(&P->A);
    free(P);
}

/*
 * Partition the matrix with the reference's acgsymcsrmatrix_partition
 * (acg/symcsrmatrix.c:685) for a given row->part map and report, for part
 * p, the sizes and index arrays that define the local ordering and the
 * halo pattern (acg/graph.c:813-1446, :1898-1981).  Two-call protocol:
 * with all array pointers NULL only the sizes are returned in sz[]:
 *   sz = {nprows, nownedrows, ninnerrows, nborderrows, nghostrows,
 *         nrecipients, sendsize, nsenders, recvsize, fnpnzs, onpnzs}
 */
int ref_partition_part(int n, int64_t nnz, const int *row, const int *col,
                       const double *a, int nparts, const int *rowparts, int p,
                       int64_t *sz, int *nzrows,
                       int *recipients, int *sendcounts, int *sendbufidx,
                       int *senders, int *recvcounts, int *recvbufidx,
                       int64_t *frowptr, int *fcolidx, double *fa,
                       int64_t *orowptr, int *ocolidx, double *oa)
{
    struct acgsymcsrmatrix A;
    int err = acgsymcsrmatrix_init_real_double(&A, n, nnz, 0, row, col, a);
    if (err) return err;
    struct acgsymcsrmatrix *parts = calloc((size_t) nparts, sizeof(*parts));
    err = acgsymcsrmatrix_partition(&A, nparts, rowparts, parts, 0);
    if (err) return err;
    struct acgsymcsrmatrix *Ap = &parts[p];
    err = acgsymcsrmatrix_dsymv_init(Ap, 0.0);
    if (err) return err;
    struct acghalo halo;
    err = acgsymcsrmatrix_halo(Ap, &halo);
    if (err) return err;
    sz[0] = Ap->nprows; sz[1] = Ap->nownedrows; sz[2] = Ap->ninnerrows;
    sz[3] = Ap->nborderrows; sz[4] = Ap->nghostrows;
    sz[5] = halo.nrecipients; sz[6] = halo.sendsize;
    sz[7] = halo.nsenders; sz[8] = halo.recvsize;
    sz[9] = Ap->fnpnzs; sz[10] = Ap->onpnzs;
    if (nzrows) {
        for (int i = 0; i < Ap->nprows; i++) nzrows[i] = Ap->nzrows[i];
        /* Before the parts are scattered to processes every neighbourrank is
         * still 0 (acg/graph.c sets ranks when it distributes subgraphs); the
         * neighbour's identity at this point is its part number,
         * acg/graph.h:288 neighbourpart.  Report that. */
        for (int i = 0; i < halo.nrecipients; i++) { recipients[i] = Ap->graph->neighbours[i].neighbourpart; sendcounts[i] = halo.sendcounts[i]; }
        for (int i = 0; i < halo.sendsize; i++) sendbufidx[i] = halo.sendbufidx[i];
        for (int i = 0; i < halo.nsenders; i++) { senders[i] = Ap->graph->neighbours[i].neighbourpart; recvcounts[i] = halo.recvcounts[i]; }
        for (int i = 0; i < halo.recvsize; i++) recvbufidx[i] = halo.recvbufidx[i];
        memcpy(frowptr, Ap->frowptr, ((size_t) Ap->nprows + 1) * sizeof(int64_t));
        memcpy(fcolidx, Ap->fcolidx, (size_t) Ap->fnpnzs * sizeof(int));
        memcpy(fa, Ap->fa, (size_t) Ap->fnpnzs * sizeof(double));
        memcpy(orowptr, Ap->orowptr, ((size_t) (Ap->nborderrows + Ap->nghostrows) + 1) * sizeof(int64_t));
        memcpy(ocolidx, Ap->ocolidx, (size_t) Ap->onpnzs * sizeof(int));
        memcpy(oa, Ap->oa, (size_t) Ap->onpnzs * sizeof(double));
    }
    acghalo_free(&halo);
    for (int q = 0; q < nparts; q++) acgsymcsrmatrix_free(&parts[q]);
    free(parts);
    acgsymcsrmatrix_free(&A);
    return 0;
}
