set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py 8584 6 > gpurun_out/r2i_fwd.log 2>&1; tail -1 gpurun_out/r2i_fwd.log
PROF_TAPE_GB=60 timeout 300 python tools/prof_case.py 8584 12 > gpurun_out/r2i_prof.log 2>&1; tail -1 gpurun_out/r2i_prof.log
timeout 600 python -m pytest tests/test_ros_adj_gpu.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -2 gpurun_out/r2i_pytest.log
