# usage: bash tools/run_scale.sh "2 4"
mkdir -p gpurun_out
for N in $1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "N=$N exit $?"; tail -2 gpurun_out/scale_n$N.err | cut -c1-300
  python - "$N" <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/scale_n{N}.json').read().strip().splitlines()[-1])
    print(f"N={N}: value {d['value']:.1f} it/s e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.2f} spmv {d['roofline']['ms_per_launch']:.4f} ms launches {d['gpu_launches']} resid {d['residual_after_step']:.6f}")
except Exception as e: print("parse fail", e)
PY
done
