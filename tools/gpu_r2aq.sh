# round 2, final evidence (3/3): ncu --set full of the kernels added in the second half of round 2, summarised on the box
set -x
mkdir -p gpurun_out /tmp/rep
timeout 300 python tools/erk_adj_case.py 296 adj > gpurun_out/r2aq_cases.log 2>&1
timeout 300 python tools/erk_adj_case.py 296 dense >> gpurun_out/r2aq_cases.log 2>&1
timeout 300 python tools/sdirk_lock_case.py 296 direct 6 >> gpurun_out/r2aq_cases.log 2>&1
timeout 300 python tools/sdirk_lock_case.py 296 trunc 6 >> gpurun_out/r2aq_cases.log 2>&1
timeout 300 python tools/sdirk_lock_case.py 296 newton_adj 6 >> gpurun_out/r2aq_cases.log 2>&1; cat gpurun_out/r2aq_cases.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_erk_adj_swe" -c 1 -o /tmp/rep/erk_adj python tools/erk_adj_case.py 296 adj > gpurun_out/r2aq_ncu_erk_adj.log 2>&1
python tools/ncu_summary.py /tmp/rep/erk_adj.ncu-rep > gpurun_out/r2aq_erk_adj_summary.txt 2>&1; python tools/ncu_lines.py /tmp/rep/erk_adj.ncu-rep k_erk_adj_swe 25 >> gpurun_out/r2aq_erk_adj_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_tlm_lock" -c 1 -o /tmp/rep/lock python tools/sdirk_lock_case.py 296 direct 6 > gpurun_out/r2aq_ncu_lock.log 2>&1
python tools/ncu_summary.py /tmp/rep/lock.ncu-rep > gpurun_out/r2aq_sdirk_tlm_lock_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_adj_newton" -c 1 -o /tmp/rep/newton python tools/sdirk_lock_case.py 296 newton_adj 6 > gpurun_out/r2aq_ncu_newton.log 2>&1
python tools/ncu_summary.py /tmp/rep/newton.ncu-rep > gpurun_out/r2aq_sdirk_adj_newton_summary.txt 2>&1
tail -22 gpurun_out/r2aq_erk_adj_summary.txt
du -sh gpurun_out
