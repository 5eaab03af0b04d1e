# Round-2 call 2 on EIGHT B200s (gpurun --gpus 8 --timeout 900 -- 'bash tools/run_r02_c2.sh'):
# every loop back-end on 8 / 2 ranks against the oracle, then A/B lines at 8 ranks in one launch per
# workload (tools/ab.py): strong-scaled 27pt-224 and the weak-scaled 27pt-448 (BASELINE config 4).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
set -x
ACGB200_TEST_ALL_BACKENDS=1 ACGB200_TEST_EXPERIMENTAL=1 timeout 500 python -m pytest tests/test_multirank.py -m gpu -q -k test_multi_gpu 2>&1 | tail -15 | tee gpurun_out/c2_pytest_multigpu.log
timeout 300 $TR --master-port 29541 tools/ab.py --workload 27pt-224 --tag c2 --solvers pipelined --variants base,pdl,onekernel,onekernel_split,unified2,unfused,nccl,nccl_graph 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/c2_ab_224_n8_pipelined.log
timeout 200 $TR --master-port 29542 tools/ab.py --workload 27pt-224 --tag c2 --solvers classic --variants base,unified2,nccl 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/c2_ab_224_n8_classic.log
timeout 300 $TR --master-port 29543 tools/ab.py --workload 27pt-448 --tag c2 --solvers pipelined,classic --variants base,onekernel 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/c2_ab_448_n8.log
timeout 200 $TR --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c2_bench_n8.json 2> gpurun_out/c2_bench_n8.err
tail -c 1500 gpurun_out/c2_bench_n8.json; tail -3 gpurun_out/c2_bench_n8.err
