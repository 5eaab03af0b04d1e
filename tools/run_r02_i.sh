# Round-2 call I on ONE B200 (last GPU seconds of the round): the slice kernel with exception rows on the parts of
# partitions, every slice shape again, and the C3 line of the plain kernel after the change.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 70 python -u -m pytest tests/test_gpu_partitioned_single.py tests/test_gpu_parity.py -m gpu -q -x -k "border_ghost or slices" > gpurun_out/i_pytest.log 2>&1
tail -3 gpurun_out/i_pytest.log
timeout 60 python tools/ab.py --workload 27pt-224 --tag i --solvers pipelined --steps 3 --warmup 1 --variants base 2>&1 | tee gpurun_out/i_ab_224.log
