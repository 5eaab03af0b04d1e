/*
 * acgb200-mpirun -- starts N ranks of a program linked against the MPI shim
 * (compat/mpi/mpishim.c) on this node:
 *
 *     acgb200-mpirun -n 8 ./acg-cuda A.mtx --comm nccl --solver acg-pipelined ...
 *
 * Exports ACGB200_MPI_RANK / ACGB200_MPI_SIZE / ACGB200_MPI_JOB to each child (and
 * LOCAL_RANK, for programs that pick their GPU from it), waits for all of them, and ends
 * the others if one fails (MPI_Abort, a crash).  Exit status: 0 if every rank returned 0,
 * otherwise the first failing rank's status.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

int main(int argc, char **argv)
{
    int n = 1, i = 1;
    while (i < argc) {
        if ((!strcmp(argv[i], "-n") || !strcmp(argv[i], "-np")) && i + 1 < argc) { n = atoi(argv[i + 1]); i += 2; }
        else if (!strcmp(argv[i], "--")) { i++; break; }
        else break;
    }
    if (n < 1 || n > 64 || i >= argc) {
        fprintf(stderr, "usage: %s -n N [--] program [arguments]   (1 <= N <= 64, one node)\n", argv[0]);
        return 2;
    }
    pid_t *pid = calloc((size_t) n, sizeof(*pid));
    if (!pid) return 2;
    char job[32], size[16];
    snprintf(job, sizeof(job), "%ld", (long) getpid());
    snprintf(size, sizeof(size), "%d", n);
    for (int r = 0; r < n; r++) {
        pid[r] = fork();
        if (pid[r] < 0) { perror("fork"); for (int q = 0; q < r; q++) kill(pid[q], SIGTERM); return 2; }
        if (pid[r] == 0) {
            char rank[16];
            snprintf(rank, sizeof(rank), "%d", r);
            setenv("ACGB200_MPI_RANK", rank, 1);
            setenv("ACGB200_MPI_SIZE", size, 1);
            setenv("ACGB200_MPI_JOB", job, 1);
            setenv("LOCAL_RANK", rank, 1);
            execvp(argv[i], &argv[i]);
            fprintf(stderr, "%s: cannot execute %s: %s\n", argv[0], argv[i], strerror(errno));
            _exit(127);
        }
    }
    int left = n, status = 0;
    while (left > 0) {
        int st = 0;
        const pid_t p = wait(&st);
        if (p < 0) { if (errno == EINTR) continue; break; }
        left--;
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        if (code != 0 && status == 0) {
            status = code;
            for (int r = 0; r < n; r++) if (pid[r] != p) kill(pid[r], SIGTERM);     /* the job is over */
        }
    }
    free(pid);
    return status;
}
