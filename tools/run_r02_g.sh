# Round-2 final call on ONE B200: the whole GPU suite, smoke(), and the default bench line, on the tree as committed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/g_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/g_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
tail -c 1200 gpurun_out/g_bench_n1.json; tail -3 gpurun_out/g_bench_n1.err
