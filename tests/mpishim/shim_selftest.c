/* Self-test of compat/mpi (the MPI stand-in): point-to-point in both directions at once, messages
 * larger than the socket buffers, tags out of order, every collective the reference's host layer
 * uses, MPI_IN_PLACE, communicator duplication.  Prints "ok" on rank 0; any mismatch aborts. */
#include <mpi.h>
#include <math.h>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "rank %d: check failed at line %d: %s\n", rank, __LINE__, #c); MPI_Abort(MPI_COMM_WORLD, 3); } } while (0)

int main(int argc, char **argv)
{
    int rank = 0, size = 1, prov = 0;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_FUNNELED, &prov);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    const int next = (rank + 1) % size, prev = (rank + size - 1) % size;

    /* head-to-head exchange of 64 MB with both neighbours (blocking sends on both sides) */
    const int big = 8 << 20;
    double *sb = malloc((size_t) big * sizeof(double)), *rb = malloc((size_t) big * sizeof(double));
    for (int i = 0; i < big; i++) sb[i] = rank + 1e-6 * i;
    if (size > 1) {
        MPI_Send(sb, big, MPI_DOUBLE, next, 7, MPI_COMM_WORLD);
        MPI_Recv(rb, big, MPI_DOUBLE, prev, 7, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        CHECK(rb[0] == prev && rb[big - 1] == prev + 1e-6 * (big - 1));
    }
    /* tags out of order; non-blocking both ways */
    int a = 100 + rank, b = 200 + rank, ra = -1, rbb = -1;
    MPI_Request rq[4];
    MPI_Isend(&a, 1, MPI_INT, next, 1, MPI_COMM_WORLD, &rq[0]);
    MPI_Isend(&b, 1, MPI_INT, next, 2, MPI_COMM_WORLD, &rq[1]);
    MPI_Irecv(&rbb, 1, MPI_INT, prev, 2, MPI_COMM_WORLD, &rq[2]);
    MPI_Irecv(&ra, 1, MPI_INT, prev, 1, MPI_COMM_WORLD, &rq[3]);
    MPI_Waitall(4, rq, MPI_STATUSES_IGNORE);
    CHECK(ra == 100 + prev && rbb == 200 + prev);

    /* collectives */
    int64_t v = rank + 1, tot = 0;
    MPI_Allreduce(&v, &tot, 1, MPI_INT64_T, MPI_SUM, MPI_COMM_WORLD);
    CHECK(tot == (int64_t) size * (size + 1) / 2);
    double d[2] = { 1.0 / (rank + 1), -(double) rank };
    MPI_Allreduce(MPI_IN_PLACE, d, 2, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
    CHECK(d[0] == 1.0 && d[1] == 0.0);
    _Bool flag = rank == size - 1;
    MPI_Allreduce(MPI_IN_PLACE, &flag, 1, MPI_C_BOOL, MPI_LOR, MPI_COMM_WORLD);
    CHECK(flag);
    int loc[2] = { rank == 0 ? 5 : 9, rank };
    MPI_Allreduce(MPI_IN_PLACE, loc, 1, MPI_2INT, MPI_MAXLOC, MPI_COMM_WORLD);
    CHECK(loc[0] == (size > 1 ? 9 : 5) && loc[1] == (size > 1 ? 1 : 0));
    double s = rank + 0.5, rs = 0;
    MPI_Reduce(&s, &rs, 1, MPI_DOUBLE, MPI_SUM, size - 1, MPI_COMM_WORLD);
    if (rank == size - 1) CHECK(fabs(rs - (size * (size - 1) / 2.0 + 0.5 * size)) < 1e-12);
    int64_t pre = -1, mine = rank + 1;
    MPI_Exscan(&mine, &pre, 1, MPI_INT64_T, MPI_SUM, MPI_COMM_WORLD);
    if (rank > 0) CHECK(pre == (int64_t) rank * (rank + 1) / 2);
    char name[64] = "";
    if (rank == 0) strcpy(name, "from-root");
    MPI_Bcast(name, 64, MPI_CHAR, 0, MPI_COMM_WORLD);
    CHECK(!strcmp(name, "from-root"));
    /* scatterv / gatherv with ragged counts */
    int *cnt = malloc((size_t) size * sizeof(int)), *dsp = malloc((size_t) size * sizeof(int)), total = 0;
    for (int r = 0; r < size; r++) { cnt[r] = r + 1; dsp[r] = total; total += cnt[r]; }
    int *all = malloc((size_t) total * sizeof(int)), *part = malloc((size_t) (rank + 1) * sizeof(int));
    if (rank == 0) for (int i = 0; i < total; i++) all[i] = 1000 + i;
    MPI_Scatterv(all, cnt, dsp, MPI_INT, part, rank + 1, MPI_INT, 0, MPI_COMM_WORLD);
    for (int i = 0; i <= rank; i++) { CHECK(part[i] == 1000 + dsp[rank] + i); part[i] *= 2; }
    memset(all, 0, (size_t) total * sizeof(int));
    MPI_Gatherv(part, rank + 1, MPI_INT, all, cnt, dsp, MPI_INT, 0, MPI_COMM_WORLD);
    if (rank == 0) for (int i = 0; i < total; i++) CHECK(all[i] == 2 * (1000 + i));
    int64_t four[4] = { rank, 2 * rank, 3 * rank, 4 * rank }, *g4 = malloc((size_t) size * 4 * sizeof(int64_t));
    MPI_Gather(four, 4, MPI_INT64_T, g4, 4, MPI_INT64_T, 0, MPI_COMM_WORLD);
    if (rank == 0) for (int r = 0; r < size; r++) CHECK(g4[4 * r + 3] == 4 * r);
    /* a duplicated communicator does not see COMM_WORLD's traffic */
    MPI_Comm dup, node;
    MPI_Comm_dup(MPI_COMM_WORLD, &dup);
    MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, &node);
    int one = rank, got = -1, nr = -1;
    MPI_Comm_rank(node, &nr);
    CHECK(nr == rank);
    if (size > 1) {
        MPI_Send(&one, 1, MPI_INT, next, 5, dup);
        one = 1000 + rank;
        MPI_Send(&one, 1, MPI_INT, next, 5, MPI_COMM_WORLD);
        MPI_Recv(&got, 1, MPI_INT, prev, 5, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
        CHECK(got == 1000 + prev);
        MPI_Recv(&got, 1, MPI_INT, prev, 5, dup, MPI_STATUS_IGNORE);
        CHECK(got == prev);
    }
    MPI_Comm_free(&dup); MPI_Comm_free(&node);
    /* derived contiguous type */
    MPI_Datatype pair;
    MPI_Type_contiguous(3, MPI_DOUBLE, &pair); MPI_Type_commit(&pair);
    double tri[3] = { rank, rank + 0.25, rank + 0.5 }, rtri[3];
    MPI_Request q2[2];
    MPI_Irecv(rtri, 1, pair, prev, 9, MPI_COMM_WORLD, &q2[0]);
    MPI_Isend(tri, 1, pair, next, 9, MPI_COMM_WORLD, &q2[1]);
    MPI_Waitall(2, q2, MPI_STATUSES_IGNORE);
    CHECK(rtri[2] == prev + 0.5);
    MPI_Type_free(&pair);
    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) printf("ok %d ranks\n", size);
    free(sb); free(rb); free(cnt); free(dsp); free(all); free(part); free(g4);
    MPI_Finalize();
    return 0;
}
