mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --solver classic --no-cpu-baseline > gpurun_out/final_bench_n1_classic.json 2>> gpurun_out/final_bench_n1.err
timeout 300 python bench.py --workload 7pt-256 --no-cpu-baseline > gpurun_out/final_bench_7pt.json 2>> gpurun_out/final_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/final_ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_tiles -s 30 -c 2 -o gpurun_out/final_prof_spmv -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/final_ncu2.log 2>&1; echo "ncu2 exit $?"
python - <<'PY'
import json
for f in ("final_bench_n1","final_bench_n1_classic","final_bench_7pt"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f spmv %.4f ms frac %.3f launches %d" % (d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['gpu_launches']), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, "fail", e)
PY
