"""CPU only: what a row partition of a BASELINE stencil matrix looks like to the SpMV plan.

    python tools/partition_study.py --n 224 --parts 8 --partition metis,block

For every part of the partition (METIS through acgsymcsrmatrix_partition_rows, as the reference does,
acg/symcsrmatrix.c:656; or the geometric blocks bench.py uses): owned / border / ghost rows, nonzeros,
and which share of the interior rows the pattern slices (slices.c) would take, with how many exception
rows -- the host-side plan acgsolvercuda_init builds, computed without a device.  METIS numbers rows
inside a part by global index, so the offsets col - row of a part are irregular wherever the part's
surface is; this measures how much of the index-free path survives that.
One JSON object per partition kind on stdout (+ a table on stderr).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402


def study_part(ab, A):
    A.dsymv_init(0.0)
    c = A.c
    cover_hi = c.borderrowoffset if (c.nghostrows > 0) else c.nownedrows
    rp, ci = A.frowptr, A.fcolidx
    s = ab.slices_host(rp, ci, cover_hi=int(cover_hi))
    out = dict(owned=int(c.nownedrows), border=int(c.nborderrows), ghost=int(c.nghostrows), fnnz=int(c.fnpnzs),
               onnz=int(c.onpnzs), interior=int(cover_hi), slice_rows=int(s["rows"]), slice_exc=int(s["nexc"]),
               slice_npat=int(s["npat"]), slice_lpad=int(s["lpad"]))
    out["slice_share_of_interior"] = out["slice_rows"] / max(out["interior"], 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--kind", type=int, default=27, choices=[7, 27])
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--partition", default="metis,block")
    ap.add_argument("--kway", action="store_true")
    args = ap.parse_args()
    import acg_b200 as ab
    from acg_b200 import dist as abdist
    N, P = args.n, args.parts
    for kind in args.partition.split(","):
        rec = dict(matrix=f"{args.kind}pt-{N}", parts=P, partition=kind)
        t0 = time.time()
        parts = []
        if kind == "block":
            for p in range(P):
                px, py, pz = abdist.grid_factors(P)
                parts.append(ab.SymCsrMatrix.stencil_part(args.kind, N, N, N, px, py, pz, p))
            rec["build_s"] = time.time() - t0
        else:
            # a whole matrix for the partitioner: generate, then re-enter through the COO constructor (as bench.make_matrix)
            G = ab.SymCsrMatrix.stencil_part(args.kind, N, N, N, 1, 1, 1, 0)
            rp = G.rowptr
            rows = np.repeat(np.arange(G.c.nprows, dtype=np.int32), np.diff(rp).astype(np.int64))
            cols, vals, n = G.colidx.copy(), G.a.copy(), int(G.c.nprows)
            G.free()
            A = ab.SymCsrMatrix.init_real_double(n, rows, cols, vals)
            del rows, cols, vals
            rec["generate_s"] = time.time() - t0
            t1 = time.time()
            rowparts, cut = A.partition_rows(P, kway=args.kway, seed=0)
            rec["metis_s"] = time.time() - t1
            rec["edge_cut"] = int(cut)
            t1 = time.time()
            parts = A.partition(P, rowparts)
            rec["split_s"] = time.time() - t1
            A.free()
        per = []
        for p, m in enumerate(parts):
            per.append(study_part(ab, m))
            m.free()
        rec["per_part"] = per
        tot_owned = sum(x["owned"] for x in per)
        rec["rows_max_over_mean"] = max(x["owned"] for x in per) / (tot_owned / P)
        rec["nnz_max_over_mean"] = max(x["fnnz"] + x["onnz"] for x in per) / (sum(x["fnnz"] + x["onnz"] for x in per) / P)
        rec["ghost_max"] = max(x["ghost"] for x in per)
        rec["ghost_total"] = sum(x["ghost"] for x in per)
        rec["border_share"] = sum(x["border"] for x in per) / tot_owned
        rec["slice_share_of_all_rows"] = sum(x["slice_rows"] for x in per) / tot_owned
        rec["slice_share_of_interior_min"] = min(x["slice_share_of_interior"] for x in per)
        rec["total_s"] = time.time() - t0
        print(json.dumps(rec), flush=True)
        print(f"# {rec['matrix']} x{P} {kind}: rows max/mean {rec['rows_max_over_mean']:.3f}, nnz max/mean {rec['nnz_max_over_mean']:.3f}, "
              f"ghosts max {rec['ghost_max']} total {rec['ghost_total']}, border rows {100 * rec['border_share']:.1f} %, "
              f"rows in slices {100 * rec['slice_share_of_all_rows']:.1f} % (min over parts of the interior share "
              f"{100 * rec['slice_share_of_interior_min']:.1f} %), {rec['total_s']:.1f} s"
              + (f", METIS {rec['metis_s']:.1f} s, cut {rec['edge_cut']}" if "metis_s" in rec else ""), file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
