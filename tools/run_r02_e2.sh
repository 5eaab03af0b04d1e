# Round-2 call E2 on TWO B200s: the multi-GPU tests call E's 400 s budget did not finish (its output was lost with
# the timeout): the unmodified driver as two processes under the MPI stand-in, and the default loop back-end of
# tests/test_multirank.py.  Output goes straight to files so that a timeout keeps what was printed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 150 python -u -m pytest tests/test_reference_driver.py -m gpu -v -k "several_gpus" > gpurun_out/e2_pytest_driver.log 2>&1
tail -8 gpurun_out/e2_pytest_driver.log
timeout 150 python -u -m pytest tests/test_multirank.py -m gpu -v -k "multi_gpu and 27pt" > gpurun_out/e2_pytest_multigpu.log 2>&1
tail -8 gpurun_out/e2_pytest_multigpu.log
