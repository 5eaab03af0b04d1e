source tools/run_variants.sh
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29544 tests/_dist_worker.py --mode gpu --matrix 27pt --size 32 --partition block 2>&1 | grep -E "^\[gpu|FAIL|Error|exitcode" | head
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 2 > gpurun_out/var_n8.json 2> gpurun_out/var_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/var_n8.json').read().strip().splitlines()[-1])
    print(f"n8: value {d['value']:.1f} it/s e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.2f} spmv {d['roofline']['ms_per_launch']:.4f} upd {d['roofline']['update_ms_per_iteration']:.4f} ms launches {d['gpu_launches']} resid {d['residual_after_step']:.6f}", flush=True)
except Exception as e:
    print("parse fail", e); print(open('gpurun_out/var_n8.err').read()[-1500:])
PY
