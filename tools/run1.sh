set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
free -g | head -2; nproc
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py parity > gpurun_out/parity.log 2>&1; echo "parity exit $?"
tail -40 gpurun_out/parity.log
timeout 600 compute-sanitizer --tool memcheck python - > gpurun_out/sanitizer.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, acg_b200 as ab
from acg_b200 import matgen as mg
n, r, c, v = mg.stencil3d_27pt(10)
A = ab.SymCsrMatrix.init_real_double(n, r, c, v).dsymv_init(0.0)
cg = ab.SolverCuda(A); b = A.vector(); b.x[:] = 1
for m in ("solvempi", "solve_pipelined"):
    x = A.vector(); print(m, getattr(cg, m)(b, x, maxits=30, residualrtol=1e-8, warmup=1), cg.c.niterations)
PY
echo "sanitizer exit $?"; tail -8 gpurun_out/sanitizer.log
N=128 timeout 600 python tools/gpu_check.py time > gpurun_out/time128.log 2>&1; echo "time128 exit $?"; cat gpurun_out/time128.log
N=224 timeout 900 python tools/gpu_check.py time > gpurun_out/time224.log 2>&1; echo "time224 exit $?"; cat gpurun_out/time224.log
