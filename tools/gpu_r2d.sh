# round 2, call d: bench with the tensor-memory forward kernel + tape hint, ncu of the new kernel
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1200 gpurun_out/r2d_bench.json
PROF_TAPE_GB=60 timeout 300 python tools/prof_case.py 8584 12 > gpurun_out/r2d_prof_plain.log 2>&1; tail -1 gpurun_out/r2d_prof_plain.log
PROF_TAPE_GB=60 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_fwd_cell_tm" -c 1 -o gpurun_out/prof_r2d_fwd_tm python tools/prof_case.py 8584 12 > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
timeout 1500 python -m pytest tests/test_ros_adj_gpu.py tests/test_units_gpu.py tests/test_small_models_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
