mkdir -p gpurun_out
run() { # name N env...
  name=$1; N=$2; shift 2
  if [ "$N" = "1" ]; then
    env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  else
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  fi
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{name}.json').read().strip().splitlines()[-1])
    print(f"{name}: value {d['value']:.1f} it/s e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.2f} spmv {d['roofline']['ms_per_launch']:.4f} upd {d['roofline']['update_ms_per_iteration']:.4f} ms launches {d['gpu_launches']} resid {d['residual_after_step']:.6f}", flush=True)
except Exception as e:
    print(name, "parse fail", e); print(open(f'gpurun_out/var_{name}.err').read()[-1500:])
PY
}
