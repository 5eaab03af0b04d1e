# Round-2 call B on ONE B200 (gpurun --timeout 1100 -- 'bash tools/run_r02_b.sh'):
# first hardware run of the pattern-slice SpMV (spmv_slices_kernel) and of the device-side full-storage
# expansion (expand.cu): parity tests, a sweep of the slice kernel's shape at C3, the 7-point and R-MAT
# workloads, set-up times, the whole GPU suite, the bench line, and ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | tee gpurun_out/b_box.log
set -x
timeout 400 python -m pytest tests/test_gpu_expand.py tests/test_gpu_parity.py -m gpu -q -k "expand or identical or golden_full or one_based or without_full or slices or spmv_matches or cg_matches or pdl or medium" 2>&1 | tail -30 | tee gpurun_out/b_pytest_new.log
timeout 420 python tools/ab.py --workload 27pt-224 --tag b --solvers pipelined --steps 3 --warmup 1 --variants base,noslices,s9,s9p,s4,s4p,s3p,s5p,s5,s14,s14p,s9p_t256,s9_t256,s5p_t256,s9p_c4,s9p_c3,s9_c8 2>&1 | tee gpurun_out/b_ab_224.log
timeout 240 python tools/ab.py --workload 7pt-256 --tag b --solvers classic --steps 3 --warmup 1 --variants base,noslices,s7,s7p,s4p,s14p,s7_t256 --opt s7_t256:slice_ub=7,slice_pf=0,slice_threads=256 2>&1 | tee gpurun_out/b_ab_7pt.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err
tail -c 2500 gpurun_out/b_bench_n1.json; tail -5 gpurun_out/b_bench_n1.err
timeout 240 python tools/setup_time.py --workload 27pt-224 2>&1 | tail -3 | tee gpurun_out/b_setup_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_slices -s 30 -c 2 -o gpurun_out/b_ncu_slices -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/b_ncu_slices.log 2>&1
tail -3 gpurun_out/b_ncu_slices.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
timeout 400 python tools/ab.py --workload rmat-20M --tag b --solvers classic,pipelined --steps 2 --warmup 1 --variants base,med64,med32 2>&1 | tee gpurun_out/b_ab_rmat20m.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tiles -s 6 -c 1 -o gpurun_out/b_ncu_rmat_tiles -f python bench.py --workload rmat-20M --steps 1 --warmup 1 --iters 10 --no-cpu-baseline > gpurun_out/b_ncu_rmat.log 2>&1
tail -3 gpurun_out/b_ncu_rmat.log
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/b_pytest_gpu.log
