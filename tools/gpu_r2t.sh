# round 2, call t: full GPU suite + the driver's default bench with the two-warps-per-cell backward kernel + ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_gpu.log; tail -5 gpurun_out/r2t_pytest_gpu.log
( time timeout 1200 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err ) 2> gpurun_out/r2t_time_default.txt; tail -c 600 gpurun_out/r2t_bench_default.json; tail -3 gpurun_out/r2t_time_default.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_bench_launches.csv python bench.py --cells 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-side > gpurun_out/r2t_ncu_bench.log 2>&1; tail -2 gpurun_out/r2t_ncu_bench.log
