source tools/run_variants.sh
ACGB200_VERBOSE=1 timeout 600 python -m pytest tests/test_multirank.py -q -m gpu -x 2>&1 | tail -5
run n2_p2p_graph 2 ACGB200_P2P=1 ACGB200_GRAPH=1 ACGB200_VERBOSE=1
run n2_p2p_nograph 2 ACGB200_P2P=1 ACGB200_GRAPH=0
run n2_nccl_graph 2 ACGB200_P2P=0 ACGB200_GRAPH=1
run n2_nccl_nograph 2 ACGB200_P2P=0 ACGB200_GRAPH=0
run n2_p2p_graph_classic 2 ACGB200_P2P=1 ACGB200_GRAPH=1 BENCH_SOLVER=classic
run n2_nccl_graph_classic 2 ACGB200_P2P=0 ACGB200_GRAPH=1 BENCH_SOLVER=classic
grep -h "peer-memory" gpurun_out/var_n2_p2p_graph.err | head -2
