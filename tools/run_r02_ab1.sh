# Round-2 opening run on ONE B200 (gpurun --timeout 1500 -- 'bash tools/run_r02_ab1.sh'):
# default GPU suite, then the opt-in experiments of round 1 that never ran on hardware,
# then A/B bench lines for each of them.  Every line lands in gpurun_out/var_<name>.json.
cd "$(dirname "$0")/.."
source tools/run_variants.sh
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ACGB200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pdl or compressed or fused or medium" 2>&1 | tail -5
set +x
run base 1
run oldgrid 1 ACGB200_BLAS1_CTAS=4
run pdl 1 ACGB200_PDL=1
run compress 1 ACGB200_SPMV_COMPRESS=1
run onekernel 1 ACGB200_PCG_FUSED=1
run onekernel_pdl 1 ACGB200_PCG_FUSED=1 ACGB200_PDL=1
run onekernel_compress 1 ACGB200_PCG_FUSED=1 ACGB200_SPMV_COMPRESS=1
run classic 1 BENCH_SOLVER=classic
run classic_pdl 1 BENCH_SOLVER=classic ACGB200_PDL=1
# one rank's share of the 8-GPU problem without any exchange: what the small size alone costs
for v in "s112 X=1" "s112_pdl ACGB200_PDL=1" "s112_oldgrid ACGB200_BLAS1_CTAS=4" "s112_onekernel ACGB200_PCG_FUSED=1"; do
  set -- $v; name=$1; shift
  env "$@" timeout 300 python bench.py --workload 27pt-112 --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{name}.json').read().strip().splitlines()[-1])
    print(f"{name}: value {d['value']:.1f} it/s ms/iter {d['ms_per_step']/d['config']['iters_per_step']:.4f} spmv {d['roofline']['ms_per_launch']:.4f} upd {d['roofline']['update_ms_per_iteration']:.4f}", flush=True)
except Exception as e:
    print(name, "parse fail", e); print(open(f'gpurun_out/var_{name}.err').read()[-1500:])
PY
done
# power-law input (config 5): plain plan against the warp-per-row kernel for medium rows
for v in "rmat2m X=1" "rmat2m_med128 ACGB200_SPMV_MEDIUM=128" "rmat2m_med256 ACGB200_SPMV_MEDIUM=256"; do
  set -- $v; name=$1; shift
  env "$@" timeout 600 python bench.py --workload rmat-2M --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{name}.json').read().strip().splitlines()[-1])
    print(f"{name}: value {d['value']:.1f} it/s spmv {d['roofline']['ms_per_launch']:.4f} ms min-traffic {d['roofline']['achieved_min_traffic_gbs']:.0f} GB/s", flush=True)
except Exception as e:
    print(name, "parse fail", e); print(open(f'gpurun_out/var_{name}.err').read()[-1500:])
PY
done
