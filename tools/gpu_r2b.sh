# round 2, call b: ROS_TLM lock-step kernel (ICNTRL(12) = 1 preset + general pivoting pass)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ros_tlm_gpu.py -m gpu -x -q --durations=10 > gpurun_out/r2b_pytest_tlm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tlm.log; tail -25 gpurun_out/r2b_pytest_tlm.log
