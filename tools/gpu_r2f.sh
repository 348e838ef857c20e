# round 2, call f: full GPU suite + parity campaign with the tensor-memory forward kernel and the lock-step ROS_TLM kernel
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log; tail -3 gpurun_out/r2f_pytest_gpu.log
timeout 1200 python tools/parity_campaign.py 128 > gpurun_out/r2f_campaign.log 2>&1; tail -2 gpurun_out/r2f_campaign.log
