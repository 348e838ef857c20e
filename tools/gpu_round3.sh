set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1p.log; tail -3 gpurun_out/pytest_gpu_r1p.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1p.log 2>&1; tail -1 gpurun_out/smoke_r1p.log
timeout 900 python bench.py > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; tail -c 400 gpurun_out/bench_r1p.json
timeout 900 python tools/parity_campaign.py 128 > gpurun_out/campaign_r1p.log 2>&1; tail -2 gpurun_out/campaign_r1p.log
for w in cbm4_rk_adj cbm4_sdirk_adj cbm4_sdirk_tlm; do timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/side3_$w.json 2> gpurun_out/side3_$w.err; tail -c 300 gpurun_out/side3_$w.json; done
