"""Worker for the multi-process tests (launched by torch.distributed.run).

mode cpu : gloo only.  Exercises the N>1 host logic end to end: every rank
           partitions the same synthetic matrix, keeps its part, exchanges
           ghost values with its neighbours following its acghalo pattern
           (gloo send/recv) and runs a distributed CG whose local arithmetic is
           numpy (test-only code); rank 0 compares with the single-rank oracle.
mode gpu : the product path: NCCL communicator inside libacgb200, device SpMV /
           halo / allreduce; rank 0 compares with the single-rank oracle.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if os.environ.get("ACGB200_TEST_HOSTSIM"):
    # host-logic run on the device stand-in of tests/hostsim (CPU test-suite only): the binding
    # is pointed at the stand-in before it loads anything
    import acg_b200.api as _api                # noqa: E402
    _api._LIBPATH = os.environ["ACGB200_TEST_HOSTSIM"]
import acg_b200 as ab                      # noqa: E402
from acg_b200 import dist as abdist        # noqa: E402
from acg_b200 import matgen as mg          # noqa: E402
from oracle import Oracle                  # noqa: E402


def halo_exchange_gloo(h, xl):
    reqs = []
    recvbufs = []
    for j, q in enumerate(h["senders"]):
        buf = torch.empty(int(h["recvcounts"][j]), dtype=torch.float64)
        recvbufs.append(buf)
        reqs.append(dist.irecv(buf, src=int(q)))
    for i, q in enumerate(h["recipients"]):
        seg = torch.from_numpy(xl[h["sendbufidx"][h["sdispls"][i]:h["sdispls"][i] + h["sendcounts"][i]]].copy())
        reqs.append(dist.isend(seg, dst=int(q)))
    for r in reqs:
        r.wait()
    for j in range(len(h["senders"])):
        idx = h["recvbufidx"][h["rdispls"][j]:h["rdispls"][j] + h["recvcounts"][j]]
        xl[idx] = recvbufs[j].numpy()


def local_matvec(m, h, xl):
    halo_exchange_gloo(h, xl)
    no = m.c.nownedrows
    y = np.zeros(no)
    rp, ci, va = m.frowptr, m.fcolidx, m.fa
    y[:] = np.add.reduceat(va * xl[ci], rp[:-1][:no])[:no] if len(va) else 0.0
    empty = np.diff(rp[:no + 1]) == 0
    y[empty] = 0.0
    b0 = m.c.borderrowoffset
    orp = m.orowptr
    for i in range(m.c.nborderrows):
        k0, k1 = orp[i], orp[i + 1]
        if k1 > k0:
            y[b0 + i] += m.oa[k0:k1] @ xl[b0 + m.ocolidx[k0:k1]]
    return y


# exchange back-ends of the CG loop (options of acgb200_set_option on top of the defaults)
BACKEND_OPTIONS = {"p2p-fused": {}, "p2p-unfused": {"p2p_fuse": 0}, "nccl": {"p2p": 0},
                   "nccl-graph": {"p2p": 0, "graph": 2}, "nccl-serial-reduce": {"p2p": 0, "redstream": 0},
                   "tiles-only": {"spmv_slices": 0}, "pdl": {"pdl": 1},
                   "watchdog": {}}


def allsum(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--matrix", default="27pt")
    ap.add_argument("--size", type=int, default=10)
    ap.add_argument("--partition", default="block")
    ap.add_argument("--maxits", type=int, default=200)
    ap.add_argument("--rtol", type=float, default=1e-9)
    ap.add_argument("--backends", default="",
                    help="gpu mode: comma-separated loop back-ends to run one after the other in this launch "
                         "(default: whatever the environment selects)")
    args = ap.parse_args()
    rank, world, _ = abdist.init_process(backend="gloo")
    N = args.size
    if args.matrix == "27pt":
        n, r, c, v = mg.stencil3d_27pt(N)
    elif args.matrix == "7pt":
        n, r, c, v = mg.laplace3d_7pt(N)
    else:
        n, r, c, v = mg.rmat_spd(N, 12 * N, seed=4)
    if args.partition == "block" and args.matrix != "rmat":
        rowparts = abdist.block_partition(N, N, N, *abdist.grid_factors(world))
    elif args.partition == "metis":
        rowparts = "metis"
    elif args.partition == "random":
        rowparts = np.random.default_rng(3).integers(0, world, n).astype(np.int32)
    else:
        rowparts = (np.arange(n) * world // n).astype(np.int32)
    if args.partition == "file":
        # ingest path: rank 0 writes aCG's binary Matrix Market file, then every rank
        # streams it and keeps its own row block (acgb200_mtx_read_part) -- nobody
        # partitions the whole matrix
        from acg_b200 import mtxio
        path = os.path.join(tempfile.gettempdir(), f"acgb200_dist_{os.environ.get('MASTER_PORT', '0')}.mtx")
        if rank == 0:
            mtxio.write_symmetric(path, n, r, c, v, binary=True)
        dist.barrier()
        m = abdist.local_part_from_file(path, "rows", rank, world)
        dist.barrier()
        if rank == 0:
            os.remove(path)
    else:
        m = abdist.local_part(n, r, c, v, rowparts, rank, world)
    no = m.c.nownedrows
    rng = np.random.default_rng(11)
    bglob = rng.standard_normal(n)
    maxits, rtol = args.maxits, args.rtol
    failures = []

    if args.mode == "cpu":
        h = m.halo()
        b = bglob[m.nzrows[:no]]
        x = np.zeros(m.c.nprows)
        rvec = b.copy()
        p = np.zeros(m.c.nprows); p[:no] = rvec
        rr = allsum(rvec @ rvec); r0 = np.sqrt(rr)
        its = 0
        for k in range(maxits):
            t = local_matvec(m, h, p)
            alpha = rr / allsum(p[:no] @ t)
            x[:no] += alpha * p[:no]
            rvec -= alpha * t
            rr_new = allsum(rvec @ rvec)
            its += 1
            if rtol > 0 and np.sqrt(rr_new) < rtol * r0:
                break
            p[:no] = rvec + (rr_new / rr) * p[:no]
            rr = rr_new
        xloc, niter = x[:no], its
        methods = [("cpu-classic", xloc, niter)]
    else:
        comm = abdist.nccl_comm(rank, world)
        assert comm.size() == world and comm.rank() == rank
        b = m.vector(); b.x[:no] = bglob[m.nzrows[:no]]
        methods = []
        defaults = {"p2p": 1, "p2p_fuse": 1, "graph": 1, "redstream": 1, "pdl": 0, "spmv_slices": 1}
        for be in (args.backends.split(",") if args.backends else [""]):
            if be:
                for key, val in {**defaults, **BACKEND_OPTIONS[be]}.items():
                    ab.set_option(key, val)
            tag = f"{be}:" if be else ""
            cg = ab.SolverCuda(m, comm)              # collective: the exchange is set up per solver
            if be == "watchdog":
                # rank 1 stops after 5 iterations and never publishes again: the other ranks' kernels
                # must give up (ACGB200_P2P_TIMEOUT_MS) and the solve must return an error, not hang
                x = m.vector()
                try:
                    code = cg.solvempi(b, x, maxits=5 if rank == 1 else 12)
                    outcome = f"returned {code}"
                except ab.AcgError as e:
                    outcome = f"error {e.code}/{e.errcode}"
                want = "returned 0" if rank == 1 else "error 4/702"      # ACG_ERR_CUDA / cudaErrorLaunchTimeout
                print(f"[gpu world={world}] watchdog rank {rank}: {outcome} {'OK' if outcome == want else 'FAIL'}", flush=True)
                if outcome != want:
                    failures.append(f"watchdog rank {rank}: {outcome}")
                cg.free()
                continue
            for meth in ("solvempi", "solve_pipelined"):
                x = m.vector()
                code = getattr(cg, meth)(b, x, maxits=maxits, residualrtol=rtol, warmup=2)
                if code != 0:
                    failures.append(f"{tag}{meth}: status {code}")
                methods.append((tag + meth, x.x[:no].copy(), cg.c.niterations))
            # fixed iteration counts, tolerances off
            x = m.vector()
            code = cg.solvempi(b, x, maxits=7)
            methods.append((tag + "solvempi-7its", x.x[:no].copy(), cg.c.niterations))
            x = m.vector()
            code = cg.solve_pipelined(b, x, maxits=9)
            methods.append((tag + "solve_pipelined-9its", x.x[:no].copy(), cg.c.niterations))
            cg.free()
        comm.destroy()

    # gather solutions on rank 0 and compare with the single-rank oracle
    gathered = [None] * world
    dist.all_gather_object(gathered, (m.nzrows[:no].copy(), [(name, xl, it) for name, xl, it in methods]))
    if rank == 0:
        O = Oracle()
        csr = O.full_csr(n, r, c, v)
        for mi, (name, _, _) in enumerate(methods):
            xg = np.zeros(n)
            its = set()
            for rows, ms in gathered:
                xg[rows] = ms[mi][1]
                its.add(ms[mi][2])
            if name.endswith("7its"):
                want = O.cg(csr, bglob, maxits=7)
            elif name.endswith("9its"):
                want = O.cg_pipelined(csr, bglob, maxits=9)
            elif name.endswith("solve_pipelined"):
                want = O.cg_pipelined(csr, bglob, maxits=maxits, rtol=rtol)
            else:
                want = O.cg(csr, bglob, maxits=maxits, rtol=rtol)
            err = np.abs(xg - want["x"]).max() / np.abs(want["x"]).max()
            ok = its == {want["niterations"]} and err < 1e-9
            print(f"[{args.mode} world={world} {args.matrix}-{N} {args.partition}] {name}: its {sorted(its)} oracle {want['niterations']} xerr {err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
            if not ok:
                failures.append(name)
    flag = torch.tensor([len(failures)], dtype=torch.int64)
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag[0]) else 0)


if __name__ == "__main__":
    main()
