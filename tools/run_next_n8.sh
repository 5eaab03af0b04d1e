# NOT RUN YET -- the 8-GPU work round 2 had no budget for, as one script:
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/run_next_n8.sh'        (about 15 box-minutes = 120 of the budget)
# Every step writes straight into gpurun_out/ (no pipes through tail: a step killed by its timeout keeps its log).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
# 1. parity of every loop back-end between N GPUs, one log per back-end
for be in p2p-fused p2p-unfused nccl nccl-graph nccl-serial-reduce tiles-only; do
  timeout 150 $TR --master-port 29541 tests/_dist_worker.py --mode gpu --matrix 27pt --size 48 --partition block --backends $be \
      > gpurun_out/n8_parity_$be.log 2>&1
  echo "backend $be exit $?" >> gpurun_out/n8_parity_summary.log
done
# 2. the exception-row slice kernel between N GPUs against tiles only (its first timing), both solvers
timeout 300 $TR --master-port 29542 tools/ab.py --workload 27pt-224 --tag n8 --variants base,noslices,unfused,nccl \
    --solvers pipelined,classic > gpurun_out/n8_ab_224.log 2>&1
# 3. the scaling line and BASELINE config 4 (27-pt 448^3 over 8 GPUs, block and METIS rows)
timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err
timeout 400 $TR --master-port 29544 bench.py --gpus $N --workload 27pt-448 --steps 5 --warmup 3 \
    > gpurun_out/n8_bench_448.json 2> gpurun_out/n8_bench_448.err
timeout 400 $TR --master-port 29545 bench.py --gpus $N --workload 27pt-224 --partition metis --steps 5 --warmup 3 \
    > gpurun_out/n8_bench_224_metis.json 2> gpurun_out/n8_bench_224_metis.err
# 4. the unmodified reference driver on N ranks through the MPI shim
timeout 200 python -m pytest tests/test_reference_driver.py -m gpu -q -k several_gpus > gpurun_out/n8_pytest_driver.log 2>&1
tail -3 gpurun_out/n8_parity_summary.log gpurun_out/n8_ab_224.log gpurun_out/n8_pytest_driver.log
tail -c 1500 gpurun_out/n8_bench.json
