# round 2, call e: pipeline mode 3 (expand beside the backward sweep) A/B
set -x
mkdir -p gpurun_out
for m in 0 3; do
FATODE_PIPELINE=$m timeout 900 python bench.py --cells 65536 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_p$m.json 2> gpurun_out/r2e_bench_p$m.err; python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench_p$m.json').read().strip().splitlines()[-1])
print('mode $m', d['value'], d['ms_per_step'], d['roofline']['forward_ms_per_step'], d['roofline']['expand_ms_per_step'], d['roofline']['adjoint_ms_per_step'], d['roofline']['launches'])
"
done
FATODE_PIPELINE=3 timeout 600 python -m pytest tests/test_ros_adj_gpu.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -3 gpurun_out/r2e_pytest.log
