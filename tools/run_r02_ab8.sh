# Round-2 8-GPU A/B (gpurun --gpus 8 --timeout 900 -- 'bash tools/run_r02_ab8.sh'):
# 2-rank correctness of the opt-in back-ends first, then strong-scaling lines for the default path, the old BLAS-1 grid and PDL.
cd "$(dirname "$0")/.."
source tools/run_variants.sh
# every loop back-end in ONE launch (one solver after the other), 8 ranks
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29544 tests/_dist_worker.py --mode gpu --matrix 27pt --size 32 --partition block --backends p2p-fused,p2p-unfused,nccl,nccl-graph,one-kernel,one-kernel-split,all-unified,two-kernel-unified,pdl,one-kernel-pdl 2>&1 | grep -E "^\[gpu|FAIL|Error|exitcode" | head -40
run n8_base 8
run n8_oldgrid 8 ACGB200_BLAS1_CTAS=4
run n8_pdl 8 ACGB200_PDL=1
run n8_onekernel 8 ACGB200_PCG_FUSED=1
run n8_onekernel_pdl 8 ACGB200_PCG_FUSED=1 ACGB200_PDL=1
run n8_onekernel_split 8 ACGB200_PCG_FUSED=1 ACGB200_P2P_UNIFIED=0
run n8_classic 8 BENCH_SOLVER=classic
run n8_classic_unified 8 BENCH_SOLVER=classic ACGB200_P2P_UNIFIED=2
run n8_twokernel_unified 8 ACGB200_P2P_UNIFIED=2
