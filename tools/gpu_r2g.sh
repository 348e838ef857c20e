# round 2, call g: correctly rounded cos / root kernels on the device — parity vs oracle and golden, cost
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_portable_math.py tests/test_ros_adj_gpu.py tests/test_ros_tlm_gpu.py tests/test_small_models_gpu.py tests/test_rk.py tests/test_sdirk.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --cells 65536 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-side > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('bench', d['value'], d['ms_per_step'], d['roofline']['forward_ms_per_step'], d['roofline']['expand_ms_per_step'], d['roofline']['adjoint_ms_per_step'])
"
