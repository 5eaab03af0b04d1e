# Round-2 call H on ONE B200: ncu of the final default slice kernel (<9,128>, 48 registers) and the launch list of a bench run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 100 ncu --set full --clock-control none --import-source on -k regex:spmv_slices -s 30 -c 2 -o gpurun_out/h_ncu_slices -f python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline > gpurun_out/h_ncu_slices.log 2>&1
tail -2 gpurun_out/h_ncu_slices.log
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_launches.log 2>&1
tail -2 gpurun_out/h_launches.log
