# Round-2 call J on ONE B200 (the round's last GPU seconds): the branch-free exception path of the slice kernel
# on the parts of partitions, and every slice shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 45 python -u -m pytest tests/test_gpu_partitioned_single.py tests/test_gpu_parity.py -m gpu -q -x -k "border_ghost or slices" > gpurun_out/j_pytest.log 2>&1
tail -3 gpurun_out/j_pytest.log
