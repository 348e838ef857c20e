set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1l.log
tail -3 gpurun_out/pytest_gpu_r1l.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1l.log 2>&1; tail -2 gpurun_out/smoke_r1l.log
timeout 900 python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -c 600 gpurun_out/bench_r1l.json
for w in cbm4_sdirk_tlm swe_erk swe_erk_tlm; do timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/side_$w.json 2> gpurun_out/side_$w.err; tail -c 900 gpurun_out/side_$w.json; done
# ncu: launch list + full capture of the new kernels (small cases)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_erk_swe -c 1 -o gpurun_out/prof_r1l_erk python tools/erk_case.py 296 > gpurun_out/ncu_erk.log 2>&1; tail -2 gpurun_out/ncu_erk.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_tlm_cell|k_sdirk_tlm_prep" -c 2 -o gpurun_out/prof_r1l_sdirk_tlm python tools/sdirk_tlm_case.py 296 > gpurun_out/ncu_sdtlm.log 2>&1; tail -2 gpurun_out/ncu_sdtlm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --cells 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-300
