source tools/run_variants.sh
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
run n1_graph 1 ACGB200_GRAPH=1
run n1_nograph 1 ACGB200_GRAPH=0
run n1_classic_graph 1 ACGB200_GRAPH=1 BENCH_SOLVER=classic
run n4_base 4 ACGB200_GRAPH=0 ACGB200_REDSTREAM=0
run n4_red 4 ACGB200_GRAPH=0 ACGB200_REDSTREAM=1
run n4_red_graph 4 ACGB200_GRAPH=1 ACGB200_REDSTREAM=1
run n4_red_graph_cta9 4 ACGB200_GRAPH=1 ACGB200_REDSTREAM=1 ACGB200_SPMV_MAX_CTAS=9
run n4_red_graph_cta8 4 ACGB200_GRAPH=1 ACGB200_REDSTREAM=1 ACGB200_SPMV_MAX_CTAS=8
run n2_red_graph 2 ACGB200_GRAPH=1 ACGB200_REDSTREAM=1
