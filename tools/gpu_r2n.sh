# round 2, call n: pair kernel with the de-duplicated accumulator code
set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_case.py 18944 72 > gpurun_out/r2n_pair.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pair.log; cat gpurun_out/r2n_pair.log
timeout 600 python -m pytest tests/test_ros_adj_gpu.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -5 gpurun_out/r2n_pytest.log
