"""Multi-GPU command-line solver on top of libacgb200 -- the part of
``cuda/acg-cuda.c`` that needs MPI in the reference (read, partition, scatter,
solve, report) done with one process per GPU under ``torch.distributed``:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m acg_b200.driver A.mtx --binary --solver acg-pipelined --partition metis

Differences from the reference driver's flow (cuda/acg-cuda.c:1297-1304 read on
root, :1486-1700 partition on root, :1782 scatter): every rank streams the
binary file and keeps its own rows (acgb200_mtx_read_part), so nobody holds the
matrix -- except for ``--partition metis``, where rank 0 reads it once to run
METIS (acgsymcsrmatrix_partition_rows) and broadcasts the row->part map.  The
options carry the reference's names; the solver report is the reference's
(acgsolvercuda_fwrite).  With one process this is a plain single-GPU run.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

SOLVERS = {"acg": "solvempi", "acg-pipelined": "solve_pipelined",
           "acg-device": "solve_device", "acg-device-pipelined": "solve_device_pipelined"}


def parse_args(argv=None):
    ap = argparse.ArgumentParser(prog="acg_b200.driver", description=__doc__.split("\n\n")[0])
    ap.add_argument("A", help="symmetric Matrix Market file (matrix coordinate real symmetric)")
    ap.add_argument("--binary", action="store_true", help="aCG binary encoding (mtx2bin); required for more than one process")
    ap.add_argument("--solver", default="acg", choices=sorted(SOLVERS))
    ap.add_argument("--max-iterations", type=int, default=100)
    ap.add_argument("--residual-atol", type=float, default=0.0)
    ap.add_argument("--residual-rtol", type=float, default=1e-9)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--epsilon", type=float, default=0.0, help="shift added to the diagonal (acgsymcsrmatrix_dsymv_init)")
    ap.add_argument("--partition", default="rows",
                    help="rows (contiguous blocks of equal row count), nnz (contiguous blocks of equal nonzero count; "
                         "rank 0 reads the matrix once), metis, or a Matrix Market vector file with 1-based part numbers "
                         "(mtxpartition output, as the reference's --partition=FILE)")
    ap.add_argument("--manufactured-solution", action="store_true",
                    help="random unit x*, b = A x*, report ||x - x*|| (cuda/acg-cuda.c:1969-1979, :2376-2385)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--output-comm-matrix", metavar="PATH", help="write the halo send-count matrix (cuda/acg-cuda.c:1713-1775)")
    ap.add_argument("--output-solution", metavar="PATH", help="write x as a Matrix Market vector (text)")
    ap.add_argument("--dry-run", action="store_true", help="read, partition and report the decomposition; no solver, no GPU")
    ap.add_argument("-v", "--verbose", action="count", default=0)
    return ap.parse_args(argv)


def row_partition(args, n, rank, world):
    """Row -> part map, identical on every rank."""
    import torch.distributed as dist
    from . import dist as abdist, mtxio
    from .api import SymCsrMatrix
    if args.partition == "rows":
        return abdist.contiguous_partition(n, world)
    if args.partition in ("metis", "nnz"):
        box = [None]
        if rank == 0:
            whole = SymCsrMatrix.read_mtx(args.A, binary=True)
            box[0] = (whole.partition_rows(world, kway=False, seed=args.seed)[0] if args.partition == "metis"
                      else abdist.balanced_rows_partition(whole, world))
            whole.free()
        dist.broadcast_object_list(box, src=0)
        return np.ascontiguousarray(box[0], np.int32)
    rowparts = mtxio.read_rowparts(args.partition)
    if len(rowparts) != n:
        raise SystemExit(f"{args.partition}: expected {n} rows, found {len(rowparts)}")
    if rowparts.min() < 0 or rowparts.max() >= world:
        raise SystemExit(f"{args.partition}: part numbers must be 1..{world}")
    return rowparts


def local_rhs(args, A, n):
    """Owned entries of b (and of x* for a manufactured solution)."""
    no = A.c.nownedrows
    gidx = A.nzrows if len(A.nzrows) >= A.c.nprows else np.arange(A.c.nprows)
    if not args.manufactured_solution:
        return np.ones(no), None
    # the same global x* on every rank: random in [-1,1], normalised (cuda/acg-cuda.c:1969-1979)
    xs = np.random.default_rng(args.seed).uniform(-1.0, 1.0, n)
    xs /= np.linalg.norm(xs)
    xloc = xs[gidx]                                   # owned + ghost entries in local order
    import scipy.sparse as sp
    rp = A.frowptr[:no + 1]
    F = sp.csr_matrix((A.fa[:rp[no]], A.fcolidx[:rp[no]] - A.c.rowidxbase, rp), shape=(no, A.c.nprows))
    b = F @ xloc
    if A.c.onpnzs > 0:
        nb, off = A.c.nborderrows, A.c.borderrowoffset
        orp = A.orowptr[:nb + 1]
        # columns of the border x ghost block are rebased by -borderrowoffset (acg/symcsrmatrix.c:838)
        O = sp.csr_matrix((A.oa[:orp[nb]], A.ocolidx[:orp[nb]] - A.c.rowidxbase + off, orp), shape=(nb, A.c.nprows))
        b[off:off + nb] += O @ xloc
    return b, xloc[:no]


def main(argv=None):
    args = parse_args(argv)
    # torchrun exports OMP_NUM_THREADS=1 unless the caller set it; the ingest and the full-storage
    # expansion are OpenMP code -- give every rank its share of the cores (before libgomp loads)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["OMP_NUM_THREADS"] = str(max(1, ncpu // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))
    import torch
    import torch.distributed as dist
    from . import dist as abdist, mtxio
    from .api import SolverCuda, SymCsrMatrix, mtx_info

    rank, world, _local = abdist.init_process(backend="gloo")
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)
    if world > 1 and not args.binary:
        raise SystemExit("more than one process needs the binary encoding (--binary): parts are read with pread")
    info = mtx_info(args.A)
    n = info["nrows"]
    t0 = time.perf_counter()
    if world == 1:
        A = SymCsrMatrix.read_mtx(args.A, binary=args.binary)
    else:
        rowparts = row_partition(args, n, rank, world)
        A = SymCsrMatrix.read_mtx_part(args.A, world, rowparts, rank)
    A.dsymv_init(args.epsilon)
    if args.verbose:
        log(f"read and partitioned in {time.perf_counter() - t0:.3f} s: {n} rows, {info['nnzs']} stored nonzeros, {world} part(s)")

    M = abdist.comm_matrix(A, rank, world)
    if args.output_comm_matrix and rank == 0:
        mtxio.write_comm_matrix(args.output_comm_matrix, M)
    if args.verbose or args.dry_run:
        rows = [None] * world
        mine = (int(A.c.nownedrows), int(A.c.ninnerrows), int(A.c.nborderrows), int(A.c.nghostrows), int(A.c.fnpnzs + A.c.onpnzs))
        if world > 1:
            dist.all_gather_object(rows, mine)
        else:
            rows = [mine]
        log("part: owned interior border ghost nonzeros sends")
        for p, r in enumerate(rows):
            log(f"  {p}: {r[0]} {r[1]} {r[2]} {r[3]} {r[4]} {int(M[p].sum())}")
    if args.dry_run:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("acg_b200.driver: no CUDA device (the solver has no CPU path; --dry-run stops before it)")
    comm = abdist.nccl_comm(rank, world)
    cg = SolverCuda(A, comm)
    b = A.vector(); x = A.vector()
    bo, xstar = local_rhs(args, A, n)
    no = A.c.nownedrows
    b.x[:] = 0.0
    b.x[:no] = bo
    code = getattr(cg, SOLVERS[args.solver])(b, x, maxits=args.max_iterations, residualatol=args.residual_atol,
                                             residualrtol=args.residual_rtol, warmup=args.warmup)
    if world > 1:
        # the report of acgsolvercuda_fwritempi (acg/cgcuda.c:1948-2216): times as the maximum,
        # flop / byte / message counters as the sum over the ranks, printed by rank 0
        mine = {k: getattr(cg.c, k) for k in SolverCuda.TIMES + SolverCuda.COUNTERS}
        every = [None] * world
        dist.all_gather_object(every, mine)
        agg = {k: max(e[k] for e in every) for k in SolverCuda.TIMES}
        agg.update({k: sum(e[k] for e in every) for k in SolverCuda.COUNTERS})
        log(cg.report(aggregate=agg))
        log(f"processes: {world}")
    else:
        log(cg.report())
    if args.manufactured_solution:
        e2 = np.array([np.sum(xstar ** 2), np.sum((x.x[:no] - xstar) ** 2)])
        if world > 1:
            t = torch.from_numpy(e2)
            dist.all_reduce(t)
        log(f"initial error 2-norm: {np.sqrt(e2[0]):.15g}")
        log(f"error 2-norm: {np.sqrt(e2[1]):.15g}")
    if args.output_solution:
        gidx = (A.nzrows[:no] if len(A.nzrows) >= no else np.arange(no)).copy()
        pieces = [None] * world
        if world > 1:
            dist.all_gather_object(pieces, (gidx, x.x[:no].copy()))
        else:
            pieces = [(gidx, x.x[:no].copy())]
        if rank == 0:
            xg = np.zeros(n)
            for gi, xv in pieces:
                xg[gi] = xv
            mtxio.write_vector(args.output_solution, xg, binary=False)
    cg.free()
    comm.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if code != 0:
        log(f"acg_b200.driver: solver returned {code} (not converged)")
    return 0 if code == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
