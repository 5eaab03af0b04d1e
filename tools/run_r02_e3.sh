# Round-2 call E3 on TWO B200s: the driver test again (stdout parse fixed), and the A/B of the early system fence
# in the pipelined update kernel (the fused update costs 20 us more than the plain one at N=2, profiles/r02/e_ab_224_n2.log).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
set -x
timeout 100 python -u -m pytest tests/test_reference_driver.py -m gpu -q -k several_gpus > gpurun_out/e3_pytest_driver.log 2>&1
tail -4 gpurun_out/e3_pytest_driver.log
timeout 170 $TR --master-port 29571 tools/ab.py --workload 27pt-224 --tag e3 --solvers pipelined --steps 6 --warmup 2 --variants base,earlyfence,base,earlyfence,earlyfence_unr2,unfused 2>&1 | grep -v "^W0\|^\*\*\*" | tee gpurun_out/e3_ab_224_n2.log
