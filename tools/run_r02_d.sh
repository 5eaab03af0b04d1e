# Round-2 call D on ONE B200 (gpurun --timeout 900 -- 'bash tools/run_r02_d.sh'): the defaults after calls B/C
# (bench lines for C3, C2 and C5), one rank's share of the 8-GPU problem on one GPU (27-pt 112^3: where the
# latency of the vector-update kernel shows), merge-tile shapes, set-up times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err
tail -c 1500 gpurun_out/d_bench_n1.json; tail -3 gpurun_out/d_bench_n1.err
timeout 200 python tools/ab.py --workload 27pt-112 --tag d --solvers pipelined --steps 6 --warmup 2 --variants base,noslices,unr2,unr2_c2,pdl,unr2_pdl,oldgrid,s9_m12,s9_t64 2>&1 | tee gpurun_out/d_ab_112.log
timeout 200 python tools/ab.py --workload 27pt-112 --tag d --solvers classic --steps 6 --warmup 2 --variants base,pdl 2>&1 | tee gpurun_out/d_ab_112_classic.log
timeout 200 python tools/ab.py --workload 27pt-224 --tag d --solvers pipelined --steps 3 --warmup 1 --variants base,unr2,unr2_c2 2>&1 | tee gpurun_out/d_ab_224.log
timeout 150 python bench.py --workload 7pt-256 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/d_bench_7pt.json 2> gpurun_out/d_bench_7pt.err
tail -c 900 gpurun_out/d_bench_7pt.json
timeout 400 python tools/ab.py --workload rmat-20M --tag d --solvers pipelined --steps 2 --warmup 1 --variants base,m2048s3,m3072,m1536,m2048t256,m2048c3 --opt m2048s3:merge_stages=3 --opt m3072:merge_items=3072 --opt m1536:merge_items=1536 --opt m2048t256:merge_threads=256 --opt m2048c3:merge_max_ctas=3 2>&1 | tee gpurun_out/d_ab_rmat20m.log
timeout 240 python tools/setup_time.py --workload 27pt-224 2>&1 | tail -3 | tee gpurun_out/d_setup_time.log
timeout 300 python bench.py --workload rmat-20M --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/d_bench_rmat20m.json 2> gpurun_out/d_bench_rmat20m.err
tail -c 1200 gpurun_out/d_bench_rmat20m.json
