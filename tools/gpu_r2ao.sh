# round 2, final evidence (1/3): full GPU suite, default bench + reference arm, all entry points, ncu launch lists (small files only)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ao_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ao_pytest_gpu.log; tail -3 gpurun_out/r2ao_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2ao_bench_default.json 2> gpurun_out/r2ao_bench_default.err ) 2> gpurun_out/r2ao_time_default.txt; tail -c 300 gpurun_out/r2ao_bench_default.json; tail -3 gpurun_out/r2ao_time_default.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ao_bench_reference.json 2> gpurun_out/r2ao_bench_reference.err; tail -c 300 gpurun_out/r2ao_bench_reference.json
timeout 600 python tools/all_entry_points_case.py > gpurun_out/r2ao_all_entry_points.log 2>&1; tail -3 gpurun_out/r2ao_all_entry_points.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ao_bench_launches.csv python bench.py --cells 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-side > gpurun_out/r2ao_ncu_bench.log 2>&1; tail -c 200 gpurun_out/r2ao_ncu_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2ao_all_entry_points_launches.csv python tools/all_entry_points_case.py > gpurun_out/r2ao_ncu_all.log 2>&1; tail -1 gpurun_out/r2ao_ncu_all.log
du -sh gpurun_out
