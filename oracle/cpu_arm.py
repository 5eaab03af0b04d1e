"""CPU arm of bench.py: the reference's own CG (acg/cg.c) timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/).  bench.py runs
this file in a subprocess -- for `--impl reference` and for the `cpu_baseline` leg of the
GPU arm -- so that the OpenMP environment is set before libgomp is loaded and nothing of
the product (libacgb200.so, torch, CUDA) is in the process:

  * one thread per PHYSICAL core of the affinity mask, pinned (OMP_PLACES = one hardware
    thread per core, OMP_PROC_BIND=close): the reference's dsymv is bandwidth-bound, SMT
    siblings add nothing and unpinned threads made round-1 numbers swing 2.6x;
  * memory interleaved over the NUMA nodes when the kernel lets us (set_mempolicy), and the
    full-storage matrix arrays re-touched by the threads that stream them (ref_shim.c,
    ref_place) -- placement only, the arithmetic is the reference's;
  * the stencil matrix is generated here with numpy (b = 1, x0 = 0), not by the product.

Usage: python oracle/cpu_arm.py --kind 27pt --N 224 --iters 5 --steps 3 --warmup 1 \
           [--matrix file.npz] [--save-x out.npy]
Prints one JSON object: {"kind": "reference"|"port", "cores": C, "times": [...], ...}.
"""
import argparse
import ctypes
import glob
import json
import os
import sys
import time


def physical_cpus():
    """One logical CPU per physical core inside the affinity mask."""
    allowed = sorted(os.sched_getaffinity(0))
    seen, pick = set(), []
    for c in allowed:
        try:
            sib = open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip()
        except OSError:
            sib = str(c)
        if sib not in seen:
            seen.add(sib)
            pick.append(c)
    return pick or allowed


def interleave_memory():
    """MPOL_INTERLEAVE over all NUMA nodes for this process; returns the node count (0: not applied)."""
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    if len(nodes) < 2:
        return 0
    mask = 0
    for n in nodes:
        mask |= 1 << n
    m = ctypes.c_ulong(mask)
    libc = ctypes.CDLL(None, use_errno=True)
    r = libc.syscall(238, 3, ctypes.byref(m), ctypes.c_ulong(max(nodes) + 2))     # set_mempolicy(MPOL_INTERLEAVE)
    return len(nodes) if r == 0 else 0


def reexec_pinned():
    if os.environ.get("ACGB200_CPU_ARM_PINNED") == "1":
        return
    cpus = physical_cpus()
    if os.environ.get("ACGB200_CPU_ARM_THREADS"):
        cpus = cpus[:int(os.environ["ACGB200_CPU_ARM_THREADS"])]
    env = dict(os.environ)
    env.update(ACGB200_CPU_ARM_PINNED="1", OMP_NUM_THREADS=str(len(cpus)),
               OMP_PLACES=",".join("{%d}" % c for c in cpus), OMP_PROC_BIND="close", OMP_DYNAMIC="false",
               OMP_WAIT_POLICY="active")
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def stencil_upper(kind, N):
    """Upper triangle + diagonal of the 7/27-point matrix on an N^3 grid (diag 6/26, neighbours -1,
    lexicographic numbering): the matrix of SURVEY.md section 8(d), entries in row-major order."""
    import numpy as np
    idx = np.arange(N ** 3, dtype=np.int64)
    x, y, z = idx % N, (idx // N) % N, idx // (N * N)
    offs = [(0, 0, 0)]            # ascending column offset dx + N dy + N^2 dz > 0
    for dz in (0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if (dz, dy, dx) <= (0, 0, 0):
                    continue
                if kind == "7pt" and abs(dx) + abs(dy) + abs(dz) != 1:
                    continue
                offs.append((dx, dy, dz))
    n, nd = N ** 3, len(offs)
    present = np.empty((n, nd), dtype=bool)          # present[i, k]: row i has the entry at offset k
    for k, (dx, dy, dz) in enumerate(offs):
        present[:, k] = (x + dx >= 0) & (x + dx < N) & (y + dy >= 0) & (y + dy < N) & (z + dz < N)
    shift = np.array([dx + N * dy + N * N * dz for (dx, dy, dz) in offs], dtype=np.int64)
    value = np.array([(26.0 if kind == "27pt" else 6.0)] + [-1.0] * (nd - 1))
    # boolean indexing walks the (row, offset) grid in row-major order: rows ascending, columns ascending
    rows = np.broadcast_to(idx.astype(np.int32)[:, None], (n, nd))[present]
    cols = (idx[:, None] + shift[None, :])[present].astype(np.int32)
    vals = np.broadcast_to(value[None, :], (n, nd))[present]
    return n, np.ascontiguousarray(rows), np.ascontiguousarray(cols), np.ascontiguousarray(vals)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="27pt")
    ap.add_argument("--N", type=int, default=224)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--matrix", default=None, help="npz with n, rows, cols, vals, b (overrides --kind/--N)")
    ap.add_argument("--save-x", default=None)
    args = ap.parse_args()
    reexec_pinned()
    nodes = interleave_memory()
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import Oracle, Ref, ref_available
    t0 = time.time()
    if args.matrix:
        d = np.load(args.matrix)
        n, r, c, v, b = int(d["n"]), d["rows"], d["cols"], d["vals"], d["b"]
    else:
        n, r, c, v = stencil_upper(args.kind, args.N)
        b = np.ones(n)
    tgen = time.time() - t0
    times, res = [], None
    t0 = time.time()
    if ref_available():
        R = Ref()
        kind, cores = "reference", R.num_threads()
        h = R.setup(n, r, c, v)
        placed = R.place(h)
        tsetup = time.time() - t0
        for s in range(args.warmup + args.steps):
            last = s == args.warmup + args.steps - 1
            res = R.solve(h, b, maxits=args.iters, want_x=bool(args.save_x) and last)
            if s >= args.warmup:
                times.append(res["tsolve"])
        R.free(h)
    else:
        O = Oracle()
        kind, cores, placed = "port", O.num_threads(), 0
        csr = O.full_csr(n, r, c, v)
        tsetup = time.time() - t0
        for s in range(args.warmup + args.steps):
            t1 = time.perf_counter()
            res = O.cg(csr, b, maxits=args.iters)
            if s >= args.warmup:
                times.append(time.perf_counter() - t1)
    if args.save_x and res is not None and res.get("x") is not None:
        np.save(args.save_x, res["x"])
    print(json.dumps(dict(kind=kind, cores=cores, times=times, iters=args.iters, n=n, nnz_upper=int(len(v)),
                          niterations=res["niterations"], bnrm2=res["bnrm2"], r0nrm2=res["r0nrm2"], rnrm2=res["rnrm2"],
                          numa_interleave_nodes=nodes, matrix_placed_by_threads=int(placed),
                          omp_places=os.environ.get("OMP_PLACES", "")[:200], gen_s=tgen, setup_s=tsetup)))


if __name__ == "__main__":
    main()
