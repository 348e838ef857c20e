set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1q.log; tail -3 gpurun_out/pytest_gpu_r1q.log
timeout 900 python tools/parity_campaign.py 128 > gpurun_out/campaign_r1q.log 2>&1; tail -1 gpurun_out/campaign_r1q.log
for w in cbm4_rk_adj cbm4_sdirk_adj cbm4_sdirk_tlm; do timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/side4_$w.json 2> gpurun_out/side4_$w.err; tail -c 300 gpurun_out/side4_$w.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_adj_cell_tm|k_rk_adj_cell_tm|k_sdirk_tlm_cell_tm" -c 1 -o gpurun_out/prof_r1q_sdirk_adj_tm python tools/sdirk_adj_case.py 592 > gpurun_out/ncu_r1q_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rk_adj_cell_tm" -c 1 -o gpurun_out/prof_r1q_rk_adj_tm python tools/rk_adj_case.py 592 > gpurun_out/ncu_r1q_b.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; tail -c 300 gpurun_out/bench_r1q.json
