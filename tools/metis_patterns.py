"""CPU only: how regular is the local numbering of a METIS part of the 27-point stencil?
Counts the distinct offset patterns (col - row sequences) among the interior rows of part 0 of an 8-way METIS
partition and the runs of consecutive rows sharing one -- what the pattern slices (compress.c, slices.c) can
and cannot use.  python tools/metis_patterns.py N   (profiles/r02/k_metis_patterns.log)"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acg_b200 as ab
N=int(sys.argv[1]); P=8
G = ab.SymCsrMatrix.stencil_part(27, N, N, N, 1, 1, 1, 0)
rp = G.rowptr
rows = np.repeat(np.arange(G.c.nprows, dtype=np.int32), np.diff(rp).astype(np.int64))
cols, vals, n = G.colidx.copy(), G.a.copy(), int(G.c.nprows)
G.free()
A = ab.SymCsrMatrix.init_real_double(n, rows, cols, vals)
rowparts, cut = A.partition_rows(P, seed=0)
parts = A.partition(P, rowparts)
m = parts[0]; m.dsymv_init(0.0)
hi = int(m.c.borderrowoffset)
rp = m.frowptr[:hi+1].copy(); ci = m.fcolidx[:rp[-1]].copy()
# distinct offset patterns among interior rows, counted directly
offs = ci - np.repeat(np.arange(hi, dtype=np.int64), np.diff(rp))
from collections import Counter
cnt = Counter()
for r in range(hi):
    cnt[offs[rp[r]:rp[r+1]].tobytes()] += 1
top = sorted(cnt.values(), reverse=True)
print(f"N={N}: part 0 interior rows {hi}, distinct offset patterns {len(cnt)}, rows covered by the 303 most frequent: {sum(top[:303])} ({100*sum(top[:303])/hi:.1f} %), by 4096: {100*sum(top[:4096])/hi:.1f} %")
# runs of consecutive rows with an identical pattern
same = 0
prev=None; run=0; runs=[]
for r in range(hi):
    k = offs[rp[r]:rp[r+1]].tobytes()
    if k==prev: run+=1
    else:
        if run: runs.append(run)
        run=1; prev=k
runs.append(run)
runs=np.array(runs)
print(f"   runs of consecutive rows with one pattern: {len(runs)} runs, mean length {runs.mean():.1f}, rows in runs >= 32: {100*runs[runs>=32].sum()/hi:.1f} %")
