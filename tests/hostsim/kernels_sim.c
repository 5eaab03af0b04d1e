/*
 * kernels_sim.c -- TEST INFRASTRUCTURE ONLY.  Plain-C stand-ins for the launch
 * interface of acg_b200/csrc/kernels.cu (internal.h), one thread, one GPU, no
 * peer memory: each function does to the "device" arrays what the CUDA kernel
 * of the same name does, including the iteration control (control words,
 * parity-buffered scalars, accumulator housekeeping).  Together with
 * cuda_mock.c this lets the CPU test-suite drive the real host code of the
 * solver (cgcuda.c: set-up, warm-up, graph capture and replay, polling,
 * convergence, reports) end to end and compare it with the oracle.  It says
 * nothing about the CUDA kernels themselves -- those are tested on the B200.
 *
 * The SpMV walks the tile plan, the medium-row list and the long-row list the
 * planner produced, so a row the plan forgot (or covered twice) shows up as a
 * wrong product.
 */
#include "internal.h"
#include "hostsim.h"

#include <math.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct gate { int iter, active; };

static struct gate gate_read(const struct acgb200_ctrl *cin, const struct acgb200_devstate *st)
{
    struct gate g = { 0, 1 };
    if (cin) { g.iter = cin->iter; g.active = cin->done == 0 && cin->iter < st->maxits; }
    return g;
}

/* ---- peer-memory exchange (several processes, windows mapped through the IPC stand-in) ---- */

#define RED_IDX(ch, parity, rank) ((((ch) * 2 + (parity)) * ACGB200_MAXR + (rank)) * 2)

static double wall_ns(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e9 * (double) ts.tv_sec + (double) ts.tv_nsec;
}

/* as p2p_spin in kernels.cu: wait for *f >= seq, give up after timeout_ns and raise the sticky flag */
static void p2p_spin(const unsigned long long *f, unsigned long long seq, struct acgb200_p2pdev *P)
{
    double t0 = 0.0;
    unsigned polls = 0;
    while (__atomic_load_n(f, __ATOMIC_ACQUIRE) < seq) {
        if ((++polls & 255u) != 0) continue;
        sched_yield();
        if (__atomic_load_n(&P->timed_out, __ATOMIC_RELAXED)) return;
        if (P->timeout_ns == 0) continue;
        const double now = wall_ns();
        if (t0 == 0.0) t0 = now;
        else if (now - t0 > (double) P->timeout_ns) { __atomic_store_n(&P->timed_out, 1ull, __ATOMIC_RELEASE); return; }
    }
}

static void p2p_wait_halo(struct acgb200_p2pdev *P, unsigned long long seq)
{
    for (int j = 0; j < P->nsenders; j++) p2p_spin(P->my_hflag + P->senders[j], seq, P);
}

static void p2p_reduce(struct acgb200_p2pdev *P, int ch, int parity, unsigned long long seq, double *out)
{
    for (int r = 0; r < P->nranks; r++) p2p_spin(P->my_rflag + ch * ACGB200_MAXR + r, seq, P);
    double a = 0.0, b = 0.0;
    for (int r = 0; r < P->nranks; r++) { const double *sl = P->my_red + RED_IDX(ch, parity, r); a += sl[0]; b += sl[1]; }
    out[0] = a; out[1] = b;
}

static double p2p_sum_slot(const struct acgb200_p2pdev *P, int ch, int parity)
{
    double a = 0.0;
    for (int r = 0; r < P->nranks; r++) a += P->my_red[RED_IDX(ch, parity, r)];
    return a;
}

static void p2p_publish_red(struct acgb200_p2pdev *P, int ch, int parity, unsigned long long seq, const double *src, int count)
{
    for (int r = 0; r < P->nranks; r++) {
        double *dst = P->peer_red[r] + RED_IDX(ch, parity, P->rank);
        dst[0] = src[0];
        dst[1] = count > 1 ? src[1] : 0.0;
        __atomic_store_n(P->peer_rflag[r] + ch * ACGB200_MAXR + P->rank, seq, __ATOMIC_RELEASE);
    }
}

static void p2p_publish_halo(struct acgb200_p2pdev *P, unsigned long long seq)
{
    for (int i = 0; i < P->nrecip; i++) __atomic_store_n(P->peer_hflag[i], seq, __ATOMIC_RELEASE);
}

static void p2p_push_row(const struct acgb200_p2pdev *P, int row, int parity, double v)
{
    const int b = row - P->borderoff;
    for (int e = P->bptr[b]; e < P->bptr[b + 1]; e++) P->peer_ghost[P->bq[e]][parity][P->bdst[e]] = v;
}

int acgb200_num_sms(void) { return 148; }
void acgb200_blas1_set_ctas_per_sm(int v) { (void) v; }
void acgb200_set_pdl(int v) { (void) v; }

int acgb200_spmv_configure(struct acgb200_spmvplan *pl)
{
    pl->smem_bytes = 0;
    pl->grid = pl->ntiles > 0 ? 1 : 1;
    return 0;
}

/* expand.cu stand-in: the reference's serial fill (acg/symcsrmatrix.c:760-851) on the "device" arrays */
int acgb200_expand_device(int n, int ghost0, int border0, int nob, int64_t pnnz,
                          const int *prp, const int *pcol, const double *pa, double eps,
                          int rp_pad, int blk_pad, struct acgb200_expanded *out, cudaStream_t stream)
{
    (void) stream; (void) pnnz;
    memset(out, 0, sizeof(*out));
    int *frp = NULL, *orp = NULL;
    if (cudaMalloc((void **) &frp, ((size_t) n + 1 + (size_t) rp_pad) * sizeof(int))) return 2;
    if (cudaMalloc((void **) &orp, ((size_t) nob + 1 + (size_t) rp_pad) * sizeof(int))) return 2;
    memset(frp, 0, ((size_t) n + 1) * sizeof(int));
    memset(orp, 0, ((size_t) nob + 1) * sizeof(int));
    for (int i = 0; i < n; i++)
        for (int k = prp[i]; k < prp[i + 1]; k++) {
            const int j = pcol[k];
            if (j < ghost0) { frp[i + 1]++; if (i != j) frp[j + 1]++; }
            else if (i >= border0) orp[i - border0 + 1]++;
        }
    for (int i = 0; i < n; i++) frp[i + 1] += frp[i];
    for (int i = 0; i < nob; i++) orp[i + 1] += orp[i];
    const int fnnz = frp[n], onnz = orp[nob];
    for (int i = 1; i <= rp_pad; i++) { frp[n + i] = fnnz; orp[nob + i] = onnz; }
    int *fcol = NULL, *ocol = NULL, *cur = calloc((size_t) n + 1, sizeof(int)), *ocur = calloc((size_t) nob + 1, sizeof(int));
    double *fa = NULL, *oa = NULL;
    if (cudaMalloc((void **) &fcol, ((size_t) fnnz + (size_t) blk_pad + 1) * sizeof(int))) return 2;
    if (cudaMalloc((void **) &fa, ((size_t) fnnz + (size_t) blk_pad + 1) * sizeof(double))) return 2;
    if (cudaMalloc((void **) &ocol, ((size_t) onnz + (size_t) blk_pad + 1) * sizeof(int))) return 2;
    if (cudaMalloc((void **) &oa, ((size_t) onnz + (size_t) blk_pad + 1) * sizeof(double))) return 2;
    memset(fcol + fnnz, 0, ((size_t) blk_pad + 1) * sizeof(int)); memset(fa + fnnz, 0, ((size_t) blk_pad + 1) * sizeof(double));
    memset(ocol + onnz, 0, ((size_t) blk_pad + 1) * sizeof(int)); memset(oa + onnz, 0, ((size_t) blk_pad + 1) * sizeof(double));
    for (int i = 0; i < n; i++)
        for (int k = prp[i]; k < prp[i + 1]; k++) {
            const int j = pcol[k];
            if (j < ghost0) {
                int l = frp[i] + cur[i]++;
                fcol[l] = j; fa[l] = pa[k] + (i == j ? eps : 0.0);
                if (i != j) { l = frp[j] + cur[j]++; fcol[l] = i; fa[l] = pa[k]; }
            } else if (i >= border0) {
                const int l = orp[i - border0] + ocur[i - border0]++;
                ocol[l] = j - border0; oa[l] = pa[k];
            }
        }
    free(cur); free(ocur);
    out->d_rowptr = frp; out->d_colidx = fcol; out->d_a = fa;
    out->d_orowptr = orp; out->d_ocolidx = ocol; out->d_oa = oa;
    out->fnnz = fnnz; out->onnz = onnz;
    return 0;
}

int acgb200_slices_fill(const struct acgb200_spmvplan *pl, const int *d_rowptr, const double *d_a, cudaStream_t stream)
{
    (void) stream;
    for (int s = 0; s < pl->nslices; s++) {
        const struct acgb200_slice sl = pl->d_slices[s];
        for (int lane = 0; lane < 32; lane++) {
            const int row = sl.row0 + lane, kb = d_rowptr[row], len = d_rowptr[row + 1] - kb;
            for (int e = 0; e < sl.len; e++)
                pl->d_sval[((size_t) sl.vblk << 5) + (size_t) e * 32 + lane] = e < len ? d_a[kb + e] : 0.0;
        }
    }
    return 0;
}

/* ---- SpMV ------------------------------------------------------------------ */


/* What spmv_issue stages for one tile: the 16-byte aligned slices of values, column indices and row
 * pointers, copied with the lengths the TMA copies use -- so a slice that runs past the padded device
 * arrays is a real out-of-bounds read here (AddressSanitizer run, tools/asan_hostsim.sh) -- and
 * checked against the stage capacity the kernel reserves in shared memory.  Returns 0 if the tile
 * does not fit. */
static int stage_tile(const struct acgb200_spmvargs *a, const struct acgb200_tile *tl, double *vals, int *cols, int *rp)
{
    const struct acgb200_spmvplan *pl = a->plan;
    const int sc = (pl->nnz_cap + 8 + 3) & ~3, rc = (pl->rows_cap + 1 + 8 + 3) & ~3;
    const int row_al = tl->row_begin & ~3;
    const int nrp = (tl->row_begin + tl->nrows + 1 - row_al + 3) & ~3;
    if (tl->nnz_al > sc || nrp > rc || (tl->k_al & 3) || (tl->nnz_al & 3)) return 0;
    memcpy(vals, a->a + tl->k_al, (size_t) tl->nnz_al * sizeof(double));
    memcpy(cols, a->colidx + tl->k_al, (size_t) tl->nnz_al * sizeof(int));
    memcpy(rp, a->rowptr + row_al, (size_t) nrp * sizeof(int));
    return 1;
}

/* one row of a staged tile, indexed as the kernel indexes shared memory */
static double staged_row(const struct acgb200_tile *tl, int lr, const double *vals, const int *cols, const int *rp, const double *x)
{
    const int *rps = rp + (tl->row_begin & 3);
    const int kb = rps[lr] - tl->k_al, ke = rps[lr + 1] - tl->k_al;
    double sum = 0.0;
    for (int k = kb; k < ke; k++) sum = fma(vals[k], x[cols[k]], sum);
    return sum;
}

/* local block, plus -- fused peer-memory mode -- the border x ghost block with ghosts from the window */
static double row_product(const struct acgb200_spmvargs *a, int row, const double **xg, int iter)
{
    double sum = 0.0;
    for (int k = a->rowptr[row]; k < a->rowptr[row + 1]; k++) sum = fma(a->a[k], a->x[a->colidx[k]], sum);
    if (a->p2p && row >= a->od_rowoffset) {
        struct acgb200_p2pdev *P = (struct acgb200_p2pdev *) a->p2p;
        if (!*xg) {
            p2p_wait_halo(P, P->hbase + (unsigned long long) iter);
            *xg = P->my_ghost[iter & 1] - a->od_nrows;
        }
        const int ob = row - a->od_rowoffset;
        for (int k = a->orowptr[ob]; k < a->orowptr[ob + 1]; k++) sum = fma(a->oa[k], (*xg)[a->ocolidx[k]], sum);
    }
    return sum;
}

static void row_epilogue(const struct acgb200_spmvargs *a, int row, double sum, double *dot)
{
    if (a->mode == SPMV_R_B_AX) {
        const double v = a->b[row] - sum;
        a->y[row] = v;
        if (row < a->dotrows) *dot = fma(v, v, *dot);
    } else {
        a->y[row] = sum;
        if (a->mode == SPMV_Y_AX_DOT && row < a->dotrows) *dot = fma(a->x[row], sum, *dot);
    }
}

/* the slice kernel (spmv_slices_kernel): every covered row through the slice-major values and the
 * zero-padded offset table, cross-checked against the row's product through the CSR arrays */
static void slices_exec(const struct acgb200_spmvargs *a, int forward)
{
    const struct acgb200_spmvplan *pl = a->plan;
    const struct gate g = gate_read(a->ctrl_in, a->st);
    if (forward && a->ctrl_in) {
        *a->ctrl_out = *a->ctrl_in;
        if (g.active) {
            const int s = g.iter & 1;
            if (a->housekeeping == 1) a->st->rr_loc[s ^ 1] = 0.0;
            if (a->housekeeping == 2) { a->st->gd_loc[s ^ 1][0] = 0.0; a->st->gd_loc[s ^ 1][1] = 0.0; }
        }
    }
    if (!g.active) return;
    double dot = 0.0;
    for (int s = 0; s < pl->nslices; s++) {
        const struct acgb200_slice sl = pl->d_slices[s];
        for (int lane = 0; lane < 32; lane++) {
            const int row = sl.row0 + lane;
            const int id = pl->d_spatid[row];
            const int exc = id == (int) ACGB200_NOPATTERN;           /* exception row: columns from the index array */
            double sum = 0.0;
            int bad = (!exc && id >= pl->slice_npat) || sl.nrows != 32 || (exc && pl->slice_exc <= 0);
            const int kb = a->rowptr[row], len = a->rowptr[row + 1] - kb;
            for (int e = 0; e < sl.len && !bad; e++) {
                const double v = pl->d_sval[((size_t) sl.vblk << 5) + (size_t) e * 32 + lane];
                const int col = exc ? (e < len ? a->colidx[kb + e] : row) : row + pl->d_spatoff[(size_t) id * pl->slice_lpad + e];
                sum = fma(v, a->x[col], sum);
            }
            /* the same row through the CSR arrays, same order */
            double direct = 0.0;
            for (int k = a->rowptr[row]; k < a->rowptr[row + 1]; k++) direct = fma(a->a[k], a->x[a->colidx[k]], direct);
            row_epilogue(a, row, (!bad && sum == direct) ? sum : NAN, &dot);
        }
    }
    if (a->acc) *a->acc += dot;
}

/* warp_sum of kernels.cu on 32 "lanes": xor-shuffle tree, every lane ends with the same value */
static double warp_tree(double *v)
{
    for (int o = 16; o > 0; o >>= 1) {
        double w[32];
        for (int l = 0; l < 32; l++) w[l] = v[l] + v[l ^ o];
        memcpy(v, w, sizeof(w));
    }
    return v[0];
}

/* the merge-path kernels (spmv_merge_kernel + spmv_merge_fix_kernel), tile by tile through the slices the
 * TMA copies stage, with the kernel's summation orders; every finished row is compared with the row's
 * product through the CSR arrays (rounding-level tolerance: the orders differ for long pieces) */
static void merge_exec(const struct acgb200_spmvargs *a, int forward)
{
    const struct acgb200_spmvplan *pl = a->plan;
    const struct gate g = gate_read(a->ctrl_in, a->st);
    if (forward && a->ctrl_in) {
        *a->ctrl_out = *a->ctrl_in;
        if (g.active) {
            const int s = g.iter & 1;
            if (a->housekeeping == 1) a->st->rr_loc[s ^ 1] = 0.0;
            if (a->housekeeping == 2) { a->st->gd_loc[s ^ 1][0] = 0.0; a->st->gd_loc[s ^ 1][1] = 0.0; }
        }
    }
    if (!g.active) return;
    const int sc = (pl->merge_items + 8 + 3) & ~3, rc = (pl->merge_items + 1 + 8 + 3) & ~3;
    double *vals = malloc((size_t) sc * sizeof(double)), *prod = malloc((size_t) sc * sizeof(double));
    int *cols = malloc((size_t) sc * sizeof(int)), *rps = malloc((size_t) rc * sizeof(int));
    unsigned char *done = calloc((size_t) pl->merge_rows + 1, 1);
    double dot = 0.0;
    int bad = 0;
    for (int t = 0; t < pl->nmtiles && !bad; t++) {
        const struct acgb200_mtile tl = pl->d_mtiles[t];
        const int k_al = tl.k0 & ~3, nnz_al = (tl.k0 + tl.nnz - k_al + 3) & ~3;
        const int row_al = tl.r0 & ~3, nrp = (tl.r0 + tl.nre + 1 - row_al + 3) & ~3;
        if (nnz_al > sc || nrp > rc || tl.nnz > pl->merge_items || tl.nre > pl->merge_items) { bad = 1; break; }
        memcpy(vals, a->a + k_al, (size_t) nnz_al * sizeof(double));
        memcpy(cols, a->colidx + k_al, (size_t) nnz_al * sizeof(int));
        memcpy(rps, a->rowptr + row_al, (size_t) nrp * sizeof(int));
        const int koff = tl.k0 & 3;
        const int *rp = rps + (tl.r0 & 3);
        for (int kk = 0; kk < tl.nnz; kk++) prod[kk] = vals[koff + kk] * a->x[cols[koff + kk]];
        const int kend = tl.k0 + tl.nnz;
        const int head = tl.nre > 0 && rp[0] < tl.k0;
        if (!head) pl->d_mpart[2 * (size_t) t] = 0.0;
        for (int j = 0; j <= tl.nre; j++) {
            const int a0 = (rp[j] > tl.k0 ? rp[j] : tl.k0) - tl.k0;
            const int b0 = (j < tl.nre ? rp[j + 1] : kend) - tl.k0;
            if (b0 > tl.nnz || (b0 > a0 && a0 < 0)) { bad = 1; break; }
            double sum = 0.0;
            if (b0 - a0 > 32) {
                double lane[32];
                for (int l = 0; l < 32; l++) { lane[l] = 0.0; for (int k = a0 + l; k < b0; k += 32) lane[l] += prod[k]; }
                sum = warp_tree(lane);
            } else for (int k = a0; k < b0; k++) sum += prod[k];
            if (j == tl.nre) pl->d_mpart[2 * (size_t) t + 1] = sum;
            else if (j == 0 && head) pl->d_mpart[2 * (size_t) t] = sum;
            else {
                const int row = tl.r0 + j;
                double direct = 0.0, scale = 0.0;
                for (int k = a->rowptr[row]; k < a->rowptr[row + 1]; k++) {
                    direct = fma(a->a[k], a->x[a->colidx[k]], direct); scale += fabs(a->a[k] * a->x[a->colidx[k]]);
                }
                if (row >= pl->merge_rows || done[row]++) bad = 1;
                row_epilogue(a, row, fabs(sum - direct) <= 1e-13 * scale + 1e-300 ? sum : NAN, &dot);
            }
        }
    }
    /* rows cut by tile boundaries */
    for (int i = 0; i < pl->nsplit && !bad; i++) {
        const struct acgb200_msplit sp = pl->d_msplit[i];
        double lane[32];
        for (int l = 0; l < 32; l++) { lane[l] = 0.0; for (int t = sp.ta + l; t < sp.tb; t += 32) lane[l] += pl->d_mpart[2 * (size_t) t + 1]; }
        double sum = warp_tree(lane);
        sum += pl->d_mpart[2 * (size_t) sp.tb];
        double direct = 0.0, scale = 0.0;
        for (int k = a->rowptr[sp.row]; k < a->rowptr[sp.row + 1]; k++) {
            direct = fma(a->a[k], a->x[a->colidx[k]], direct); scale += fabs(a->a[k] * a->x[a->colidx[k]]);
        }
        if (sp.row >= pl->merge_rows || done[sp.row]++) bad = 1;
        row_epilogue(a, sp.row, fabs(sum - direct) <= 1e-13 * scale + 1e-300 ? sum : NAN, &dot);
    }
    for (int r = 0; r < pl->merge_rows; r++) if (done[r] != 1) bad = 1;      /* every row exactly once */
    if (bad) for (int r = 0; r < pl->merge_rows; r++) a->y[r] = NAN;
    free(vals); free(prod); free(cols); free(rps); free(done);
    if (a->acc) *a->acc += dot;
}

static void spmv_exec(void *p)
{
    const struct acgb200_spmvargs *a = p;
    const struct acgb200_spmvplan *pl = a->plan;
    const int slices_forward = (pl->nslices > 0 || pl->nmtiles > 0) && pl->ntiles == 0 && !a->p2p;
    if (pl->nmtiles > 0) merge_exec(a, slices_forward);
    if (pl->nslices > 0) slices_exec(a, slices_forward);
    /* the tile kernel: control word forwarding and housekeeping, then the tiles */
    if (pl->ntiles > 0 || (a->ctrl_in && !slices_forward)) {
        const struct gate g = gate_read(a->ctrl_in, a->st);
        if (a->ctrl_in) {
            *a->ctrl_out = *a->ctrl_in;
            if (g.active) {
                const int s = g.iter & 1;
                if (a->housekeeping == 1) a->st->rr_loc[s ^ 1] = 0.0;
                if (a->housekeeping == 2) { a->st->gd_loc[s ^ 1][0] = 0.0; a->st->gd_loc[s ^ 1][1] = 0.0; }
            }
        }
        if (g.active) {
            double dot = 0.0;
            const double *xg = NULL;
            const int sc = (pl->nnz_cap + 8 + 3) & ~3, rc = (pl->rows_cap + 1 + 8 + 3) & ~3;
            double *svals = malloc((size_t) sc * sizeof(double));
            int *scols = malloc((size_t) sc * sizeof(int)), *srp = malloc((size_t) rc * sizeof(int));
            for (int t = 0; t < pl->ntiles; t++) {
                const struct acgb200_tile tl = pl->d_tiles[t];
                const int staged = stage_tile(a, &tl, svals, scols, srp);
                if (a->p2p && !xg && tl.row_begin + tl.nrows > a->od_rowoffset) {
                    /* as in the kernel: wait for the neighbours before the first tile that reaches the
                     * border rows touches anything */
                    struct acgb200_p2pdev *P = (struct acgb200_p2pdev *) a->p2p;
                    p2p_wait_halo(P, P->hbase + (unsigned long long) g.iter);
                    xg = P->my_ghost[g.iter & 1] - a->od_nrows;
                }
                for (int r = tl.row_begin; r < tl.row_begin + tl.nrows; r++) {
                    /* the product through the staged slices must be the product through the arrays */
                    const double direct = row_product(a, r, &xg, g.iter);
                    double via_stage = staged ? staged_row(&tl, r - tl.row_begin, svals, scols, srp, a->x) : NAN;
                    if (staged && a->p2p && r >= a->od_rowoffset && xg) {
                        const int ob = r - a->od_rowoffset;
                        for (int k = a->orowptr[ob]; k < a->orowptr[ob + 1]; k++) via_stage = fma(a->oa[k], xg[a->ocolidx[k]], via_stage);
                    }
                    row_epilogue(a, r, (staged && via_stage == direct) ? direct : NAN, &dot);
                }
            }
            free(svals); free(scols); free(srp);
            if (a->acc) *a->acc += dot;
            if (a->p2p && a->pub_ch >= 0 && a->p2p->fuse) {
                struct acgb200_p2pdev *P = (struct acgb200_p2pdev *) a->p2p;
                p2p_publish_red(P, a->pub_ch, g.iter & 1, P->rbase + (unsigned long long) g.iter + 1ull, a->acc, 1);
            }
        }
    }
    /* medium and long rows: separate kernels, gated by the same incoming word */
    for (int pass = 0; pass < 2; pass++) {
        const int cnt = pass == 0 ? pl->nmed : pl->nlong;
        const int *rows = pass == 0 ? pl->d_medrows : pl->d_longrows;
        if (cnt <= 0) continue;
        const struct gate g = gate_read(a->ctrl_in, a->st);
        if (!g.active) continue;
        double dot = 0.0;
        const double *xg = NULL;
        if (a->p2p) {                            /* the row-list kernels wait at their start */
            struct acgb200_p2pdev *P = (struct acgb200_p2pdev *) a->p2p;
            p2p_wait_halo(P, P->hbase + (unsigned long long) g.iter);
            xg = P->my_ghost[g.iter & 1] - a->od_nrows;
        }
        for (int i = 0; i < cnt; i++) row_epilogue(a, rows[i], row_product(a, rows[i], &xg, g.iter), &dot);
        if (a->acc) *a->acc += dot;
    }
}

int acgb200_spmv_launch(const struct acgb200_spmvargs *a, cudaStream_t stream)
{
    (void) stream;
    if (a->plan->nlong > 0 && !a->plan->d_long_scratch) return 1;
    return hostsim_run_or_record(spmv_exec, a, sizeof(*a));
}

static void offdiag_exec(void *vp)
{
    const struct acgb200_offdiagargs *a = vp;
    const struct gate g = gate_read(a->ctrl_in, a->st);
    if (!g.active) return;
    const double *xg = a->x + a->rowoffset;      /* ghost tail of x: ocolidx is rebased by -borderrowoffset */
    if (a->p2p) {
        struct acgb200_p2pdev *P = (struct acgb200_p2pdev *) a->p2p;
        const int it = a->p2p_iter_override >= 0 ? a->p2p_iter_override : g.iter;
        p2p_wait_halo(P, P->hbase + (unsigned long long) it);
        xg = P->my_ghost[it & 1] - a->nrows;
    }
    double dot = 0.0;
    for (int i = 0; i < a->nrows; i++) {
        const int kb = a->orowptr[i], ke = a->orowptr[i + 1], row = a->rowoffset + i;
        double sum = 0.0;
        for (int k = kb; k < ke; k++) sum = fma(a->oa[k], xg[a->ocolidx[k]], sum);
        double v = a->y[row];
        if (kb != ke) { v = a->minus ? v - sum : v + sum; a->y[row] = v; }
        if (a->dotkind == 1) dot = fma(a->x[row], v, dot);
        else if (a->dotkind == 2) dot = fma(v, v, dot);
    }
    if (a->acc) *a->acc += dot;
}

int acgb200_offdiag_launch(const struct acgb200_offdiagargs *a, cudaStream_t stream)
{
    (void) stream;
    if (a->nrows <= 0) return 0;
    return hostsim_run_or_record(offdiag_exec, a, sizeof(*a));
}

static void post_exec(void *vp)
{
    const struct acgb200_postargs *a = vp;
    struct acgb200_p2pdev *P = a->p2p;
    int iter = a->iter_override;
    if (iter < 0) {
        const struct gate g = gate_read(a->cin, a->st);
        if (!g.active) return;
        iter = g.iter;
    }
    if (a->vec) {
        const int parity = iter & 1;
        for (int q = 0; q < P->nrecip; q++)
            for (int i = P->sdispls[q]; i < P->sdispls[q + 1]; i++)
                P->peer_ghost[q][parity][P->peer_rdispl[q] + (i - P->sdispls[q])] = a->vec[a->sendbufidx[i]];
    }
    if (a->ch >= 0) {
        const int parity = (iter + a->par_off) & 1;
        const double *src = a->redbase + parity * a->redstride;
        for (int r = 0; r < P->nranks; r++) {
            double *dst = P->peer_red[r] + RED_IDX(a->ch, parity, P->rank);
            dst[0] = src[0];
            dst[1] = a->redcount > 1 ? src[1] : 0.0;
        }
    }
    if (a->vec) p2p_publish_halo(P, P->hbase + (unsigned long long) iter);
    if (a->ch >= 0)
        for (int r = 0; r < P->nranks; r++)
            __atomic_store_n(P->peer_rflag[r] + a->ch * ACGB200_MAXR + P->rank,
                             P->rbase + (unsigned long long) (iter + a->seq_off), __ATOMIC_RELEASE);
}

int acgb200_comm_post(const struct acgb200_postargs *a, cudaStream_t stream)
{
    (void) stream;
    return hostsim_run_or_record(post_exec, a, sizeof(*a));
}

/* ---- classic CG ------------------------------------------------------------- */

struct upd_args {
    int n; struct acgb200_devstate *st; int cin, cout, multi; struct acgb200_p2pdev *P; double *wout;
    const double *q; double *z, *w, *t, *p, *r, *x;
};

static void update_r_exec(void *vp)
{
    const struct upd_args *a = vp;
    struct acgb200_devstate *st = a->st;
    const struct gate g = gate_read(&st->ctrl[a->cin], st);
    if (a->cin != a->cout) st->ctrl[a->cout] = st->ctrl[a->cin];
    if (!g.active) return;
    const int s = g.iter & 1;
    struct acgb200_p2pdev *P = a->P;
    double rr, pap;
    if (P) {
        double glob[2];
        p2p_reduce(P, 0, s, P->rbase + (unsigned long long) g.iter + 1ull, glob);
        pap = glob[0];
        rr = g.iter > 0 ? p2p_sum_slot(P, 1, s) : st->rr[0];
    } else {
        rr = a->multi ? st->rr[s] : st->rr_loc[s];
        pap = a->multi ? st->pap[s] : st->pap_loc[s];
    }
    const double alpha = rr / pap;
    double acc = 0.0;
    for (int i = 0; i < a->n; i++) {
        const double rv = fma(-alpha, a->t[i], a->r[i]);
        a->r[i] = rv;
        acc = fma(rv, rv, acc);
    }
    st->rr_loc[s ^ 1] += acc;
    if (P && P->fuse) p2p_publish_red(P, 1, s ^ 1, P->rbase + (unsigned long long) g.iter + 1ull, &st->rr_loc[s ^ 1], 1);
}

int acgb200_cg_update_r(int n, struct acgb200_devstate *st, int cin, int cout, int multi, struct acgb200_p2pdev *p2p,
                        const double *t, double *r, cudaStream_t stream)
{
    (void) stream;
    struct upd_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.st = st; a.cin = cin; a.cout = cout; a.multi = multi; a.P = p2p; a.t = (double *) t; a.r = r;
    return hostsim_run_or_record(update_r_exec, &a, sizeof(a));
}

static void update_xp_exec(void *vp)
{
    const struct upd_args *a = vp;
    struct acgb200_devstate *st = a->st;
    const struct gate g = gate_read(&st->ctrl[a->cin], st);
    const int s = g.iter & 1;
    struct acgb200_p2pdev *P = a->P;
    double rr, rrn, pap;
    if (P && g.active) {
        double glob[2];
        p2p_reduce(P, 1, s ^ 1, P->rbase + (unsigned long long) g.iter + 1ull, glob);
        rrn = glob[0];
        rr = g.iter > 0 ? p2p_sum_slot(P, 1, s) : st->rr[0];
        pap = p2p_sum_slot(P, 0, s);
    } else {
        rr = a->multi ? st->rr[s] : st->rr_loc[s];
        rrn = a->multi ? st->rr[s ^ 1] : st->rr_loc[s ^ 1];
        pap = a->multi ? st->pap[s] : st->pap_loc[s];
    }
    struct acgb200_ctrl c = st->ctrl[a->cin];
    if (g.active) {
        c.iter = g.iter + 1;
        if (P) st->rr[s ^ 1] = rrn;
        if (st->tol > 0.0 && sqrt(rrn) < st->tol) { c.done = 1; st->final_rr = rrn; }
        st->pap_loc[s ^ 1] = 0.0;
    }
    st->ctrl[a->cout] = c;
    if (!g.active) return;
    const double alpha = rr / pap, beta = rrn / rr;
    for (int i = 0; i < a->n; i++) {
        const double pv = a->p[i];
        a->x[i] = fma(alpha, pv, a->x[i]);
        const double pn = fma(beta, pv, a->r[i]);
        a->p[i] = pn;
        if (P && P->fuse && i >= P->borderoff) p2p_push_row(P, i, s ^ 1, pn);
    }
    if (P && P->fuse) p2p_publish_halo(P, P->hbase + (unsigned long long) g.iter + 1ull);
}

int acgb200_cg_update_xp(int n, struct acgb200_devstate *st, int cin, int cout, int multi, struct acgb200_p2pdev *p2p,
                         const double *r, double *p, double *x, cudaStream_t stream)
{
    (void) stream;
    struct upd_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.st = st; a.cin = cin; a.cout = cout; a.multi = multi; a.P = p2p; a.r = (double *) r; a.p = p; a.x = x;
    return hostsim_run_or_record(update_xp_exec, &a, sizeof(a));
}

/* ---- pipelined CG ------------------------------------------------------------ */

static void pcg_update_exec(void *vp)
{
    const struct upd_args *a = vp;
    struct acgb200_devstate *st = a->st;
    const struct gate g = gate_read(&st->ctrl[a->cin], st);
    const int s = g.iter & 1;
    struct acgb200_p2pdev *P = a->P;
    double gamma, delta;
    if (P && g.active && g.iter > 0) {
        double glob[2];
        p2p_reduce(P, 0, s, P->rbase + (unsigned long long) g.iter, glob);
        gamma = glob[0]; delta = glob[1];
    } else {
        gamma = a->multi ? st->gd[s][0] : st->gd_loc[s][0];
        delta = a->multi ? st->gd[s][1] : st->gd_loc[s][1];
    }
    const double gamma_prev = st->prev[s][0], alpha_prev = st->prev[s][1];
    const int conv = st->tol > 0.0 && sqrt(gamma) < st->tol;
    const double beta = gamma / gamma_prev;
    const double alpha = gamma / (delta - beta * gamma / alpha_prev);
    struct acgb200_ctrl c = st->ctrl[a->cin];
    if (g.active) {
        if (P) { st->gd[s][0] = gamma; st->gd[s][1] = delta; }
        if (conv) { c.done = 1; st->final_rr = gamma; }
        else { c.iter = g.iter + 1; st->prev[s ^ 1][0] = gamma; st->prev[s ^ 1][1] = alpha; }
    }
    st->ctrl[a->cout] = c;
    if (!g.active || conv) return;
    double g2 = 0.0, d2 = 0.0;
    /* as the kernel: border rows first in fused peer-memory mode */
    const int push = P && P->fuse, first = push ? P->borderoff : 0;
    for (int ii = 0; ii < a->n; ii++) {
        const int i = (first + ii) % (a->n > 0 ? a->n : 1);
        const double zv = fma(beta, a->z[i], a->q[i]);
        const double tv = fma(beta, a->t[i], a->w[i]);
        const double pv = fma(beta, a->p[i], a->r[i]);
        const double rv = fma(-alpha, tv, a->r[i]);
        const double wv = fma(-alpha, zv, a->w[i]);
        a->z[i] = zv; a->t[i] = tv; a->p[i] = pv;
        a->x[i] = fma(alpha, pv, a->x[i]);
        a->r[i] = rv;
        a->w[i] = wv;
        g2 = fma(rv, rv, g2);
        d2 = fma(wv, rv, d2);
        if (push && i >= first) p2p_push_row(P, i, s ^ 1, wv);
    }
    st->gd_loc[s ^ 1][0] += g2;
    st->gd_loc[s ^ 1][1] += d2;
    if (push) {
        const unsigned long long it1 = (unsigned long long) g.iter + 1ull;
        p2p_publish_red(P, 0, s ^ 1, P->rbase + it1, &st->gd_loc[s ^ 1][0], 2);
        p2p_publish_halo(P, P->hbase + it1);
    }
}

int acgb200_pcg_update(int n, struct acgb200_devstate *st, int cin, int cout, int multi, struct acgb200_p2pdev *p2p,
                       const double *q, double *z, double *w, double *t, double *p, double *r, double *x,
                       cudaStream_t stream)
{
    (void) stream;
    struct upd_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.st = st; a.cin = cin; a.cout = cout; a.multi = multi; a.P = p2p;
    a.q = q; a.z = z; a.w = w; a.t = t; a.p = p; a.r = r; a.x = x;
    return hostsim_run_or_record(pcg_update_exec, &a, sizeof(a));
}

/* ---- set-up helpers ------------------------------------------------------------ */

struct dot_args { int n; const double *x, *y; double *acc; int two; };

static void dot_exec(void *vp)
{
    const struct dot_args *a = vp;
    double g = 0.0, d = 0.0;
    for (int i = 0; i < a->n; i++) {
        if (a->two) { g = fma(a->x[i], a->x[i], g); d = fma(a->y[i], a->x[i], d); }
        else g = fma(a->x[i], a->y[i], g);
    }
    a->acc[0] += g;
    if (a->two) a->acc[1] += d;
}

int acgb200_dot(int n, const double *x, const double *y, double *acc, cudaStream_t stream)
{
    (void) stream;
    struct dot_args a = { n, x, y, acc, 0 };
    return hostsim_run_or_record(dot_exec, &a, sizeof(a));
}

int acgb200_dot2(int n, const double *r, const double *w, double *acc2, cudaStream_t stream)
{
    (void) stream;
    struct dot_args a = { n, r, w, acc2, 1 };
    return hostsim_run_or_record(dot_exec, &a, sizeof(a));
}

struct gs_args { int n; double *dst; const double *src; const int *idx; int scatter; };

static void gs_exec(void *vp)
{
    const struct gs_args *a = vp;
    if (a->scatter) for (int i = 0; i < a->n; i++) a->dst[a->idx[i]] = a->src[i];
    else for (int i = 0; i < a->n; i++) a->dst[i] = a->src[a->idx[i]];
}

int acgb200_gather(int n, double *dst, const double *src, const int *idx, cudaStream_t stream)
{
    (void) stream;
    if (n <= 0) return 0;
    struct gs_args a = { n, dst, src, idx, 0 };
    return hostsim_run_or_record(gs_exec, &a, sizeof(a));
}

int acgb200_scatter(int n, const double *src, double *dst, const int *idx, cudaStream_t stream)
{
    (void) stream;
    if (n <= 0) return 0;
    struct gs_args a = { n, dst, src, idx, 1 };
    return hostsim_run_or_record(gs_exec, &a, sizeof(a));
}

/* ---- the reference's public BLAS-1 building blocks (immediate: they are not part of any captured loop) ---- */

int acgb200_helper_axpy(int op, int n, const double *num, const double *den, const double *x, double *y, cudaStream_t stream)
{
    (void) stream;
    const double a = op == 1 ? -(*num) / (*den) : (*num) / (*den);
    for (int i = 0; i < n; i++) y[i] = op == 2 ? fma(a, y[i], x[i]) : fma(a, x[i], y[i]);
    return 0;
}

int acgb200_helper_scalars(int op, double *out0, double *out1, const double *num, const double *den, cudaStream_t stream)
{
    (void) stream;
    const double q = (*num) / (*den);
    *out0 = q;
    if (op == 0) *out1 = -q;
    return 0;
}

int acgb200_helper_pipelined(int n, const double *gamma, double *gamma_prev, const double *delta, const double *q,
                             double *p, double *r, double *t, double *x, double *z, double *w, double *alpha_prev,
                             cudaStream_t stream)
{
    (void) stream;
    const double beta = (*gamma) / (*gamma_prev);
    const double alpha = (*gamma) / ((*delta) - beta * (*gamma) / (*alpha_prev));
    for (int i = 0; i < n; i++) {
        const double zv = fma(beta, z[i], q[i]), tv = fma(beta, t[i], w[i]), pv = fma(beta, p[i], r[i]);
        z[i] = zv; t[i] = tv; p[i] = pv;
        x[i] = fma(alpha, pv, x[i]);
        r[i] = fma(-alpha, tv, r[i]);
        w[i] = fma(-alpha, zv, w[i]);
    }
    *gamma_prev = *gamma; *alpha_prev = alpha;
    return 0;
}
