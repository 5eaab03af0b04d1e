set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --solver classic --no-cpu-baseline > gpurun_out/bench_n1_classic.json 2>> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1_classic.json
timeout 600 python bench.py --workload 7pt-256 --no-cpu-baseline > gpurun_out/bench_7pt.json 2>> gpurun_out/bench_n1.err; cat gpurun_out/bench_7pt.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu1 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_tiles -s 30 -c 2 -o gpurun_out/prof_spmv -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/ncu_spmv.log 2>&1; echo "ncu2 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_update|cg_update' -s 30 -c 2 -o gpurun_out/prof_blas1 -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/ncu_blas1.log 2>&1; echo "ncu3 exit $?"
ls -la gpurun_out
