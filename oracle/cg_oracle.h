/*
 * cg_oracle.h -- CPU restatement of the aCG conjugate-gradient hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product (acg_b200/) never links, imports
 * or executes it.
 *
 * Parity status: PINNED for the classic-CG path -- every function below is
 * checked bit-for-bit (compiled with -ffp-contract=off on both sides)
 * against the reference's own C sources built into oracle/_ref/libacgref.so
 * (tests/test_oracle_pin.py, run in the build container where
 * /root/reference exists) and against fixtures in tests/golden/ generated
 * from that library (tools/make_golden.py).  The pipelined-CG recurrences
 * have no CPU implementation in the reference (they exist only as CUDA in
 * acg/cg-kernels-cuda.cu:187-221 + acg/cgcuda.c:1676-1788); for those the
 * oracle is a restatement pinned only through its agreement with the pinned
 * classic path (KAT-6) -- "parity unpinned" at the arithmetic level.
 */
#ifndef CG_ORACLE_H
#define CG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes returned in *status by the solvers (mirror enum acgerrcode
 * values of acg/error.h:49-104 that the path can produce) */
#define ORACLE_SUCCESS 0
#define ORACLE_ERR_NOT_CONVERGED 39
#define ORACLE_ERR_NOT_CONVERGED_INDEFINITE 40

/*
 * Expand an upper-triangular COO matrix (0-based, any order) into the
 * "full storage" CSR the reference multiplies with: first the row-sorted
 * packed upper CSR (acg/symcsrmatrix.c:66-131: counting sort by row,
 * stable in input order), then the mirror pass of
 * acgsymcsrmatrix_dsymv_init (acg/symcsrmatrix.c:760-812): rows visited in
 * order, each packed entry (i,j) appended to row i and, if i != j, to row
 * j; eps is added to diagonal entries.
 *
 * frowptr must hold n+1 entries; fcolidx/fa must hold 2*nnz entries
 * (upper bound).  Returns the number of full nonzeros.
 */
int64_t oracle_full_csr(
    int n, int64_t nnz, const int *rowidx, const int *colidx, const double *a,
    double eps, int64_t *frowptr, int *fcolidx, double *fa);

/* y = alpha*A*x + beta*y on full-storage CSR, restating the 4-row x 2-nnz
 * unrolled loop of acgsymcsrmatrix_dsymv (acg/symcsrmatrix.c:863-959);
 * beta is applied first as in acgvector_dscal (acg/vector.c:482-500). */
void oracle_dsymv(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    double alpha, const double *x, double beta, double *y);

/* BLAS-1 with the reference's summation order (acg/vector.c:507-653) */
double oracle_ddot(int n, const double *x, const double *y);
double oracle_dnrm2sqr(int n, const double *x);
void oracle_daxpy(int n, double a, const double *x, double *y);
void oracle_daypx(int n, double a, double *y, const double *x);

struct oracle_cg_result {
    int status;          /* ORACLE_* */
    int niterations;
    double bnrm2, r0nrm2, rnrm2;
};

/*
 * Classic CG, restating acgsolver_solve (acg/cg.c:198-386).  x holds the
 * initial guess on entry and the solution on exit.  If rnrm2hist is not
 * NULL it must hold maxits+1 doubles and receives ||r_k|| for k=0..niter.
 */
void oracle_cg(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    const double *b, double *x, int maxits,
    double residualatol, double residualrtol,
    struct oracle_cg_result *res, double *rnrm2hist);

/*
 * Pipelined (Ghysels-Vanroose) CG, restating the host loop of
 * acgsolvercuda_solve_pipelined (acg/cgcuda.c:1577-1788) with the fused
 * update of acg/cg-kernels-cuda.cu:201-214, using the CPU kernels above.
 */
void oracle_cg_pipelined(
    int n, const int64_t *frowptr, const int *fcolidx, const double *fa,
    const double *b, double *x, int maxits,
    double residualatol, double residualrtol,
    struct oracle_cg_result *res, double *rnrm2hist);

/* number of OpenMP threads the oracle will use (1 if built without) */
int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
