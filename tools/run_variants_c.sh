source tools/run_variants.sh
run n4_p2p_graph 4 ACGB200_P2P=1 ACGB200_GRAPH=1
run n4_p2p_nograph 4 ACGB200_P2P=1 ACGB200_GRAPH=0
run n4_nccl_graph 4 ACGB200_P2P=0 ACGB200_GRAPH=1
run n4_p2p_graph_classic 4 ACGB200_P2P=1 ACGB200_GRAPH=1 BENCH_SOLVER=classic
run n4_nccl_graph_classic 4 ACGB200_P2P=0 ACGB200_GRAPH=1 BENCH_SOLVER=classic
timeout 600 python -m pytest tests/test_multirank.py -q -m gpu -x 2>&1 | tail -3
