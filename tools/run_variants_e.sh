source tools/run_variants.sh
timeout 600 python -m pytest tests/test_multirank.py -q -m gpu 2>&1 | tail -3
run n2_fuse 2 ACGB200_P2P_FUSE=1
run n2_nofuse 2 ACGB200_P2P_FUSE=0
run n2_nccl 2 ACGB200_P2P=0
run n2_fuse_classic 2 ACGB200_P2P_FUSE=1 BENCH_SOLVER=classic
run n2_nofuse_classic 2 ACGB200_P2P_FUSE=0 BENCH_SOLVER=classic
