# Round-2 call E on TWO B200s (gpurun --gpus 2 --timeout 600 -- 'bash tools/run_r02_e.sh'):
# every loop back-end between two GPUs against the oracle (with pattern slices for the interior rows),
# the UNMODIFIED reference driver as two processes under the MPI stand-in, A/B lines and the bench line at N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | tee gpurun_out/e_box.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
set -x
ACGB200_TEST_ALL_BACKENDS=1 ACGB200_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_multirank.py tests/test_reference_driver.py tests/test_driver_py.py -m gpu -q -k "multi_gpu or several_gpus or solves_from_file" 2>&1 | tail -15 | tee gpurun_out/e_pytest_2gpu.log
timeout 200 $TR --master-port 29551 tools/ab.py --workload 27pt-224 --tag e --solvers pipelined --steps 4 --warmup 2 --variants base,noslices,unr2,pdl,unfused,nccl 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/e_ab_224_n2.log
timeout 120 $TR --master-port 29552 tools/ab.py --workload 27pt-224 --tag e --solvers classic --steps 4 --warmup 2 --variants base,noslices 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/e_ab_224_n2_classic.log
timeout 150 $TR --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
tail -c 1500 gpurun_out/e_bench_n2.json; tail -3 gpurun_out/e_bench_n2.err
