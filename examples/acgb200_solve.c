/*
 * acgb200_solve.c -- minimal C program on top of libacgb200 alone (no reference
 * code, no MPI, no Python): read a symmetric Matrix Market file, solve A x = b
 * with b = 1 on one B200, print the reference's solver report.
 *
 *   acgb200_solve A.mtx [--binary] [--solver acg|acg-pipelined] [--max-iterations N]
 *                       [--residual-rtol TOL] [--warmup N] [--epsilon EPS] [--print-solution]
 *
 * The same steps the reference driver takes around the solver
 * (cuda/acg-cuda.c:1297-1304 read, :1949-1967 b = 1, :2209 init, :2242-2262
 * solve, :2270 report), with the option names it uses.
 */
#include "acgb200/cgcuda.h"
#include "acgb200/error.h"
#include "acgb200/ext.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void usage(FILE *f)
{
    fprintf(f, "usage: acgb200_solve A.mtx [--binary] [--solver acg|acg-pipelined] [--max-iterations N]\n"
               "                     [--residual-rtol TOL] [--warmup N] [--epsilon EPS] [--print-solution]\n");
}

int main(int argc, char **argv)
{
    const char *path = NULL, *solver = "acg";
    int binary = 0, maxits = 100, warmup = 10, print_x = 0;
    double rtol = 1e-9, eps = 0.0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { usage(stdout); return 0; }
        else if (!strcmp(argv[i], "--binary")) binary = 1;
        else if (!strcmp(argv[i], "--print-solution")) print_x = 1;
        else if (!strcmp(argv[i], "--solver") && i + 1 < argc) solver = argv[++i];
        else if (!strcmp(argv[i], "--max-iterations") && i + 1 < argc) maxits = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--residual-rtol") && i + 1 < argc) rtol = atof(argv[++i]);
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--epsilon") && i + 1 < argc) eps = atof(argv[++i]);
        else if (argv[i][0] != '-' && !path) path = argv[i];
        else { usage(stderr); return 2; }
    }
    if (!path) { usage(stderr); return 2; }
    const int pipelined = !strcmp(solver, "acg-pipelined");
    if (!pipelined && strcmp(solver, "acg")) { usage(stderr); return 2; }

    struct acgsymcsrmatrix A;
    int err = acgb200_mtx_read(path, binary, &A);
    if (err) { fprintf(stderr, "%s: %s\n", path, acgerrcodestr(err, 0)); return 1; }
    err = acgsymcsrmatrix_dsymv_init(&A, eps);
    if (err) { fprintf(stderr, "dsymv_init: %s\n", acgerrcodestr(err, 0)); return 1; }

    struct acgvector b, x;
    if ((err = acgsymcsrmatrix_vector(&A, &b)) || (err = acgsymcsrmatrix_vector(&A, &x))) {
        fprintf(stderr, "vector: %s\n", acgerrcodestr(err, 0)); return 1;
    }
    for (acgidx_t i = 0; i < b.num_nonzeros; i++) b.x[i] = 1.0;
    acgvector_setzero(&x);

    struct acgcomm comm;
    memset(&comm, 0, sizeof(comm));
    comm.type = acgcomm_null;
    struct acgsolvercuda cg;
    err = acgsolvercuda_init(&cg, &A, NULL, NULL, &comm);
    if (err) { fprintf(stderr, "acgsolvercuda_init: %s\n", acgerrcodestr(err, 0)); return 1; }
    int errcode = 0;
    err = pipelined
        ? acgsolvercuda_solve_pipelined(&cg, &A, &b, &x, maxits, 0, 0, 0, rtol, warmup, &comm, 0, &errcode, NULL, NULL)
        : acgsolvercuda_solvempi(&cg, &A, &b, &x, maxits, 0, 0, 0, rtol, warmup, &comm, 0, &errcode, NULL, NULL, 0);
    acgsolvercuda_fwrite(stderr, &cg, 0);
    if (err) fprintf(stderr, "solver: %s\n", acgerrcodestr(err, errcode));
    if (print_x) {
        printf("%%%%MatrixMarket vector array real general\n%lld\n", (long long) x.num_nonzeros);
        for (acgidx_t i = 0; i < x.num_nonzeros; i++) printf("%.17g\n", x.x[i]);
    }
    acgsolvercuda_free(&cg);
    acgvector_free(&b); acgvector_free(&x);
    acgsymcsrmatrix_free(&A);
    return err ? 1 : 0;
}
