# Round-2 call 1 on ONE B200 (gpurun --timeout 1500 -- 'bash tools/run_r02_c1.sh'):
# the default GPU suite, the opt-in kernels of round 1 that never ran on hardware, then A/B lines
# of every variant in one process per workload (tools/ab.py), then the new bench line with the
# full-size parity leg.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|NUMA node\(s\)|Model name" 
set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/c1_pytest_gpu.log
ACGB200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pdl or compressed or fused or medium" 2>&1 | tail -25 | tee gpurun_out/c1_pytest_experimental.log
timeout 500 python tools/ab.py --workload 27pt-224 --tag c1 --variants base,oldgrid,pdl,compress,compress_pdl,onekernel,onekernel_pdl,onekernel_compress --solvers pipelined 2>&1 | tee gpurun_out/c1_ab_224_pipelined.log
timeout 300 python tools/ab.py --workload 27pt-224 --tag c1 --variants base,pdl,compress,compress_pdl --solvers classic 2>&1 | tee gpurun_out/c1_ab_224_classic.log
timeout 200 python tools/ab.py --workload 27pt-112 --tag c1 --variants base,pdl,compress,onekernel,onekernel_compress --solvers pipelined 2>&1 | tee gpurun_out/c1_ab_112.log
timeout 300 python tools/ab.py --workload rmat-2M --tag c1 --variants base,med64,med128,med256 --solvers pipelined --steps 3 2>&1 | tee gpurun_out/c1_ab_rmat2m.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err
tail -c 3000 gpurun_out/c1_bench_n1.json; tail -5 gpurun_out/c1_bench_n1.err
