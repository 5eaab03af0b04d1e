/*
 * cg_oracle.c -- CPU restatement of the aCG conjugate-gradient hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cg_oracle.h).  Plain C, no dependencies
 * beyond libm and (optionally) OpenMP.  Compile with -ffp-contract=off so
 * that the arithmetic is the literal sequence of IEEE operations written
 * here; that is what makes the bit-for-bit pin against oracle/_ref
 * meaningful.
 *
 * Every routine cites the reference lines it follows (paths relative to
 * the aCG tree).
 */
#include "cg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * acg/symcsrmatrix.c:66-131 -- COO -> row-sorted packed upper CSR.  The
 * reference counts entries per row, prefix-sums, and (if the input is not
 * already row-sorted) scatters entries in input order, so equal-row
 * entries keep their relative order.  A stable counting sort reproduces
 * both branches.
 *
 * acg/symcsrmatrix.c:760-812 -- packed -> full storage.  Row lengths of
 * the full matrix are counted (each off-diagonal packed entry contributes
 * to both its row and its column), prefix-summed, and the fill pass walks
 * packed rows in increasing order appending (i,j) to row i and (j,i) to
 * row j.  The resulting column order inside full row r is therefore: the
 * mirrored entries (i<r) in increasing i, then r's own packed entries in
 * packed order.
 */
int64_t oracle_full_csr(
    int n, int64_t nnz, const int *rowidx, const int *colidx, const double *a,
    double eps, int64_t *frowptr, int *fcolidx, double *fa)
{
    int64_t *prow = calloc((size_t)n + 1, sizeof(*prow));
    int *pcol = malloc((size_t)(nnz ? nnz : 1) * sizeof(*pcol));
    double *pval = malloc((size_t)(nnz ? nnz : 1) * sizeof(*pval));
    if (!prow || !pcol || !pval) { free(prow); free(pcol); free(pval); return -1; }

    for (int64_t k = 0; k < nnz; k++) prow[rowidx[k] + 1]++;
    for (int i = 0; i < n; i++) prow[i + 1] += prow[i];
    {
        int64_t *cursor = malloc((size_t)(n ? n : 1) * sizeof(*cursor));
        if (!cursor) { free(prow); free(pcol); free(pval); return -1; }
        memcpy(cursor, prow, (size_t)n * sizeof(*cursor));
        for (int64_t k = 0; k < nnz; k++) {
            int64_t l = cursor[rowidx[k]]++;
            pcol[l] = colidx[k];
            pval[l] = a[k];
        }
        free(cursor);
    }

    for (int i = 0; i <= n; i++) frowptr[i] = 0;
    for (int i = 0; i < n; i++) {
        for (int64_t k = prow[i]; k < prow[i + 1]; k++) {
            int j = pcol[k];
            frowptr[i + 1]++;
            if (i != j) frowptr[j + 1]++;
        }
    }
    for (int i = 0; i < n; i++) frowptr[i + 1] += frowptr[i];
    int64_t fnnz = frowptr[n];

    int64_t *cursor = malloc((size_t)(n ? n : 1) * sizeof(*cursor));
    if (!cursor) { free(prow); free(pcol); free(pval); return -1; }
    memcpy(cursor, frowptr, (size_t)n * sizeof(*cursor));
    for (int i = 0; i < n; i++) {
        for (int64_t k = prow[i]; k < prow[i + 1]; k++) {
            int j = pcol[k];
            int64_t l = cursor[i]++;
            fcolidx[l] = j;
            fa[l] = pval[k] + ((i == j) ? eps : 0.0);
            if (i != j) {
                int64_t m = cursor[j]++;
                fcolidx[m] = i;
                fa[m] = pval[k];
            }
        }
    }
    free(cursor); free(prow); free(pcol); free(pval);
    return fnnz;
}

/*
 * One row of acgsymcsrmatrix_dsymv's unrolled loop
 * (acg/symcsrmatrix.c:909-925): products are summed in pairs,
 * z += (a[k]*x[j[k]] + a[k+1]*x[j[k+1]]), with a single trailing element
 * when the row length is odd.  The 4-row / 2-row / 1-row blocking of the
 * reference only groups rows; every row's arithmetic is this function.
 */
static inline double row_sum_pairs(
    const int64_t *rp, const int *j, const double *a, const double *x, int i)
{
    int64_t lo = rp[i], hi = rp[i + 1];
    int64_t even_end = hi - (hi - lo) % 2;
    double z = 0;
    for (int64_t k = lo; k < even_end; k += 2)
        z += a[k] * x[j[k]] + a[k + 1] * x[j[k + 1]];
    for (int64_t k = even_end; k < hi; k++)
        z += a[k] * x[j[k]];
    return z;
}

void oracle_dsymv(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    double alpha, const double *x, double beta, double *y)
{
    /* acg/symcsrmatrix.c:884-887 with acg/vector.c:482-500 */
    if (beta != 1) {
        if (beta == 0) { for (int i = 0; i < n; i++) y[i] = 0; }
        else { for (int i = 0; i < n; i++) y[i] *= beta; }
    }
    /* acg/symcsrmatrix.c:903-958; rows are independent, so the OpenMP
     * schedule does not affect the result */
    #pragma omp parallel for
    for (int i = 0; i < n; i++)
        y[i] += alpha * row_sum_pairs(frowptr, fcolidx, fa, x, i);
}

/* acg/vector.c:561-593: four interleaved partial sums, remainder into the
 * first, final sum c1+c2+c3+c4 left to right */
double oracle_ddot(int n, const double *x, const double *y)
{
    double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    int m = n - n % 4;
    for (int k = 0; k < m; k += 4) {
        c1 += x[k + 0] * y[k + 0];
        c2 += x[k + 1] * y[k + 1];
        c3 += x[k + 2] * y[k + 2];
        c4 += x[k + 3] * y[k + 3];
    }
    for (int k = m; k < n; k++) c1 += x[k] * y[k];
    return c1 + c2 + c3 + c4;
}

/* acg/vector.c:631-653: same scheme with y == x */
double oracle_dnrm2sqr(int n, const double *x)
{
    return oracle_ddot(n, x, x);
}

/* acg/vector.c:507-523 */
void oracle_daxpy(int n, double a, const double *x, double *y)
{
    for (int k = 0; k < n; k++) y[k] += a * x[k];
}

/* acg/vector.c:533-550 */
void oracle_daypx(int n, double a, double *y, const double *x)
{
    for (int k = 0; k < n; k++) y[k] = a * y[k] + x[k];
}

/*
 * acg/cg.c:198-386.  Order of operations per iteration: t=Ap, pAp, alpha,
 * x+=alpha p, r-=alpha t, rr, convergence test, beta, p=beta p+r.  The
 * test compares norms (sqrt of the squared norm) with strict '<'; the
 * relative tolerance is scaled by ||r0|| once (cg.c:281).
 */
void oracle_cg(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    const double *b, double *x, int maxits,
    double residualatol, double residualrtol,
    struct oracle_cg_result *res, double *rnrm2hist)
{
    double *r = malloc((size_t)(n ? n : 1) * sizeof(*r));
    double *p = malloc((size_t)(n ? n : 1) * sizeof(*p));
    double *t = calloc((size_t)(n ? n : 1), sizeof(*t));
    res->status = ORACLE_SUCCESS;
    res->niterations = 0;
    res->bnrm2 = sqrt(oracle_dnrm2sqr(n, b));               /* cg.c:243 */
    memcpy(r, b, (size_t)n * sizeof(*r));                    /* cg.c:257 */
    oracle_dsymv(n, frowptr, fcolidx, fa, -1.0, x, 1.0, r);  /* cg.c:262 */
    memcpy(p, r, (size_t)n * sizeof(*p));                    /* cg.c:268 */
    double rr = oracle_dnrm2sqr(n, r);                       /* cg.c:275 */
    res->rnrm2 = res->r0nrm2 = sqrt(rr);
    if (rnrm2hist) rnrm2hist[0] = res->rnrm2;
    double rtol_abs = residualrtol * res->r0nrm2;            /* cg.c:281 */
    if ((residualatol > 0 && res->rnrm2 < residualatol) ||
        (rtol_abs > 0 && res->rnrm2 < rtol_abs))
        goto done;                                           /* cg.c:284-289 */

    for (int k = 0; k < maxits; k++) {
        oracle_dsymv(n, frowptr, fcolidx, fa, 1.0, p, 0.0, t);  /* cg.c:293 */
        double pAp = oracle_ddot(n, p, t);                       /* cg.c:301 */
        if (pAp == 0) { res->status = ORACLE_ERR_NOT_CONVERGED_INDEFINITE; goto done; }
        double alpha = rr / pAp;                                 /* cg.c:305 */
        oracle_daxpy(n, alpha, p, x);                            /* cg.c:315 */
        oracle_daxpy(n, -alpha, t, r);                           /* cg.c:331 */
        double rr_prev = rr;
        rr = oracle_dnrm2sqr(n, r);                              /* cg.c:338 */
        res->rnrm2 = sqrt(rr);
        if (rnrm2hist) rnrm2hist[k + 1] = res->rnrm2;
        if ((residualatol > 0 && res->rnrm2 < residualatol) ||
            (rtol_abs > 0 && res->rnrm2 < rtol_abs)) {
            res->niterations++;                                  /* cg.c:348 */
            goto done;
        }
        if (rr_prev == 0) { res->status = ORACLE_ERR_NOT_CONVERGED_INDEFINITE; goto done; }
        double beta = rr / rr_prev;                              /* cg.c:357 */
        oracle_daypx(n, beta, p, r);                             /* cg.c:361 */
        res->niterations++;
    }
    /* cg.c:375-382; note the reference tests the *scaled* relative
     * tolerance here (it multiplied residualrtol by ||r0|| in place) */
    if (!(residualatol == 0 && rtol_abs == 0))
        res->status = ORACLE_ERR_NOT_CONVERGED;
done:
    free(r); free(p); free(t);
}

/*
 * acg/cgcuda.c:1577-1788 (host loop) + acg/cg-kernels-cuda.cu:201-214
 * (fused update).  Setup: r=b-Ax0, w=Ar, z=t=p=0, alpha_prev=gamma_prev=inf
 * (cgcuda.c:1513-1522).  Iteration k: gamma=(r,r), delta=(w,r), q=Aw,
 * convergence test on sqrt(gamma) (k==0 defines ||r0||, cgcuda.c:1761),
 * then beta=gamma/gamma_prev, alpha=gamma/(delta-beta*gamma/alpha_prev),
 * z=q+beta z, t=w+beta t, p=r+beta p, x+=alpha p, r-=alpha t, w-=alpha z.
 * niterations counts completed updates (cgcuda.c:1787).
 */
void oracle_cg_pipelined(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    const double *b, double *x, int maxits,
    double residualatol, double residualrtol,
    struct oracle_cg_result *res, double *rnrm2hist)
{
    size_t m = (size_t)(n ? n : 1);
    double *r = malloc(m * sizeof(double)), *w = calloc(m, sizeof(double));
    double *q = calloc(m, sizeof(double)), *z = calloc(m, sizeof(double));
    double *t = calloc(m, sizeof(double)), *p = calloc(m, sizeof(double));
    res->status = ORACLE_SUCCESS;
    res->niterations = 0;
    res->bnrm2 = sqrt(oracle_dnrm2sqr(n, b));
    res->r0nrm2 = res->rnrm2 = INFINITY;
    memcpy(r, b, (size_t)n * sizeof(*r));
    oracle_dsymv(n, frowptr, fcolidx, fa, -1.0, x, 1.0, r);
    oracle_dsymv(n, frowptr, fcolidx, fa, 1.0, r, 0.0, w);
    double gamma_prev = INFINITY, alpha_prev = INFINITY;
    double rtol_abs = 0;
    int converged = 0;
    for (int k = 0; k < maxits; k++) {
        double gamma = oracle_dnrm2sqr(n, r);
        double delta = oracle_ddot(n, w, r);
        oracle_dsymv(n, frowptr, fcolidx, fa, 1.0, w, 0.0, q);
        res->rnrm2 = sqrt(gamma);
        if (rnrm2hist) rnrm2hist[k] = res->rnrm2;
        if (k == 0) { res->r0nrm2 = res->rnrm2; rtol_abs = residualrtol * res->r0nrm2; }
        if ((residualatol > 0 && res->rnrm2 < residualatol) ||
            (rtol_abs > 0 && res->rnrm2 < rtol_abs)) { converged = 1; break; }
        double beta = gamma / gamma_prev;
        double alpha = gamma / (delta - beta * gamma / alpha_prev);
        for (int i = 0; i < n; i++) {
            z[i] = q[i] + beta * z[i];
            t[i] = w[i] + beta * t[i];
            p[i] = r[i] + beta * p[i];
            x[i] += alpha * p[i];
            r[i] -= alpha * t[i];
            w[i] -= alpha * z[i];
        }
        gamma_prev = gamma;
        alpha_prev = alpha;
        res->niterations++;
    }
    /* cgcuda.c:1867-1874: as in the classic solver the relative tolerance
     * tested here is the scaled one (scaled at k==0, cgcuda.c:1761) */
    if (!converged && !(residualatol == 0 && (maxits > 0 ? rtol_abs : residualrtol) == 0))
        res->status = ORACLE_ERR_NOT_CONVERGED;
    free(r); free(w); free(q); free(z); free(t); free(p);
}
