# round 2, call j: the driver's two bench arms as the driver runs them (default flags, and --steps 20 --warmup 5)
set -x
mkdir -p gpurun_out
( time timeout 1700 python bench.py > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err ) 2> gpurun_out/r2j_time_default.txt; tail -c 400 gpurun_out/r2j_bench_default.json; tail -3 gpurun_out/r2j_time_default.txt
( time timeout 900 python bench.py --impl reference > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err ) 2> gpurun_out/r2j_time_ref.txt; tail -c 300 gpurun_out/r2j_bench_ref.json; tail -3 gpurun_out/r2j_time_ref.txt
