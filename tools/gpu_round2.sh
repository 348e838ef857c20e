set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sdirk_tlm_gpu.py tests/test_sdirk_adj.py tests/test_erk_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in cbm4_sdirk_tlm cbm4_sdirk_adj swe_erk swe_erk_tlm; do timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/side2_$w.json 2> gpurun_out/side2_$w.err; tail -c 700 gpurun_out/side2_$w.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_erk_swe -c 1 -o gpurun_out/prof_r1m_erk python tools/erk_case.py 296 > gpurun_out/ncu_erk2.log 2>&1; tail -2 gpurun_out/ncu_erk2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_adj_cell|k_sdirk_adj_prep" -c 2 -o gpurun_out/prof_r1m_sdirk_adj python tools/sdirk_adj_case.py 444 > gpurun_out/ncu_sdadj.log 2>&1; tail -2 gpurun_out/ncu_sdadj.log
