# round 2, call c: tensor-memory forward kernel (two warps per SM) — parity and A/B timing
set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py 8584 6 > gpurun_out/r2c_fwd_tm.log 2>&1; tail -1 gpurun_out/r2c_fwd_tm.log
FATODE_FWD_TM=0 timeout 300 python tools/prof_fwd.py 8584 6 > gpurun_out/r2c_fwd_old.log 2>&1; tail -1 gpurun_out/r2c_fwd_old.log
timeout 1500 python -m pytest tests/test_ros_adj_gpu.py tests/test_ros_tlm_gpu.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --cells 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_tm.json 2> gpurun_out/r2c_bench_tm.err; tail -c 1500 gpurun_out/r2c_bench_tm.json
