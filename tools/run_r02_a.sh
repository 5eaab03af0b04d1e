# Round-2 call A on ONE B200 (gpurun --timeout 900 -- 'bash tools/run_r02_a.sh'):
# the opt-in kernels of round 1 that never ran on hardware, the single-GPU tests of the partitioned paths,
# A/B lines of every variant in one process per workload (tools/ab.py), the bench line with the full-size
# parity leg, and the stock reference GPU solver on the same matrix.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee gpurun_out/a_box.log
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|NUMA node\(s\)|Model name" | tee -a gpurun_out/a_box.log
set -x
ACGB200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pdl or compressed or fused or medium" 2>&1 | tail -40 | tee gpurun_out/a_pytest_experimental.log
timeout 200 python -m pytest tests/test_gpu_partitioned_single.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/a_pytest_partitioned.log
timeout 300 python tools/ab.py --workload 27pt-224 --tag a --variants base,compress,onekernel,onekernel_compress,pdl,oldgrid,compress_pdl --solvers pipelined 2>&1 | tee gpurun_out/a_ab_224_pipelined.log
timeout 150 python tools/ab.py --workload 27pt-224 --tag a --variants base,compress --solvers classic 2>&1 | tee gpurun_out/a_ab_224_classic.log
timeout 300 python tools/ab.py --workload rmat-20M --tag a --variants base,med64,med256 --solvers pipelined --steps 2 --warmup 1 2>&1 | tee gpurun_out/a_ab_rmat20m.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
tail -c 2500 gpurun_out/a_bench_n1.json; tail -5 gpurun_out/a_bench_n1.err
timeout 150 python bench.py --workload 7pt-256 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/a_bench_7pt.json 2> gpurun_out/a_bench_7pt.err
tail -c 1200 gpurun_out/a_bench_7pt.json
timeout 500 python bench.py --no-cpu-baseline --with-reference-gpu --steps 3 --warmup 3 > gpurun_out/a_bench_refgpu.json 2> gpurun_out/a_bench_refgpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/a_bench_refgpu.json').read().strip().splitlines()[-1])
    print("ours", d["value"], "reference_gpu", d.get("reference_gpu"))
except Exception as e:
    print("refgpu parse fail", e); print(open('gpurun_out/a_bench_refgpu.err').read()[-1500:])
PY
