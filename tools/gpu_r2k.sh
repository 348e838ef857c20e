set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_small_models_gpu.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -30 gpurun_out/r2k_pytest.log
