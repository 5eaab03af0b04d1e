# Round-2 call C on ONE B200 (gpurun --timeout 1300 -- 'bash tools/run_r02_c.sh'):
# whole GPU suite (with the merge-path tiles' first hardware run), register-cap variants of the slice kernel,
# merge-path tile shapes on R-MAT 20 M against the row-aligned tiles, bench line, ncu captures, set-up times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee gpurun_out/c_box.log
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "merge_path or medium or ragged or residual_history" 2>&1 | tail -15 | tee gpurun_out/c_pytest_merge.log
timeout 300 python tools/ab.py --workload 27pt-224 --tag c --solvers pipelined --steps 3 --warmup 1 --variants base,s9_m10,s9_m12,s14,s14_m6,s14_m8,s9_t64,s9_t64_m20,s9_t256,s9_t256_m5,s7,s5 2>&1 | tee gpurun_out/c_ab_224.log
timeout 200 python tools/ab.py --workload 7pt-256 --tag c --solvers classic --steps 3 --warmup 1 --variants base,s7_m10,s7_m12,s7_t64,s14 2>&1 | tee gpurun_out/c_ab_7pt.log
timeout 500 python tools/ab.py --workload rmat-20M --tag c --solvers pipelined --steps 2 --warmup 1 --variants base,m512,m2048,m4096t256,m1024t256,rowtiles_med64 --opt m512:merge_items=512 --opt m2048:merge_items=2048 --opt m4096t256:merge_items=4096,merge_threads=256 --opt m1024t256:merge_threads=256 --opt rowtiles_med64:spmv_merge=0,spmv_medium=64 2>&1 | tee gpurun_out/c_ab_rmat20m.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err
tail -c 1800 gpurun_out/c_bench_n1.json; tail -5 gpurun_out/c_bench_n1.err
timeout 300 python bench.py --workload rmat-20M --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/c_bench_rmat20m.json 2> gpurun_out/c_bench_rmat20m.err
tail -c 1500 gpurun_out/c_bench_rmat20m.json
timeout 240 python tools/setup_time.py --workload 27pt-224 2>&1 | tail -3 | tee gpurun_out/c_setup_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_slices -s 30 -c 2 -o gpurun_out/c_ncu_slices -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/c_ncu_slices.log 2>&1
tail -2 gpurun_out/c_ncu_slices.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 6 -c 1 -o gpurun_out/c_ncu_rmat_merge -f python bench.py --workload rmat-20M --steps 1 --warmup 1 --iters 10 --no-cpu-baseline > gpurun_out/c_ncu_rmat.log 2>&1
tail -2 gpurun_out/c_ncu_rmat.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_launches.log 2>&1
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/c_pytest_gpu.log
