# round 2, call l: two-warps-per-cell backward sweep (k_ros_adj_pair) — parity tests, then A/B timing against k_ros_adj_cell
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ros_adj_gpu.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -15 gpurun_out/r2l_pytest.log
timeout 300 python tools/prof_case.py 18944 72 > gpurun_out/r2l_pair.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pair.log; cat gpurun_out/r2l_pair.log
FATODE_ADJ_PAIR=0 timeout 300 python tools/prof_case.py 18944 72 > gpurun_out/r2l_single.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_single.log; cat gpurun_out/r2l_single.log
