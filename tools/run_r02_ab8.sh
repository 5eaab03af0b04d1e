# Round-2 8-GPU A/B (gpurun --gpus 8 --timeout 900 -- 'bash tools/run_r02_ab8.sh'):
# strong-scaling lines for the default path, the old BLAS-1 grid and PDL.
cd "$(dirname "$0")/.."
source tools/run_variants.sh
run n8_base 8
run n8_oldgrid 8 ACGB200_BLAS1_CTAS=4
run n8_pdl 8 ACGB200_PDL=1
run n8_classic 8 BENCH_SOLVER=classic
