# round 2, call ac: ncu --set full of k_expand_tape from the HEAD binary (2368 cells, 12 h window)
set -x
mkdir -p gpurun_out
PROF_TAPE_GB=24 timeout 300 python tools/prof_case.py 2368 12 > gpurun_out/r2ac_prof_plain.log 2>&1; tail -1 gpurun_out/r2ac_prof_plain.log
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_expand_tape" -c 1 -o gpurun_out/prof_r2ac_expand python tools/prof_case.py 2368 12 > gpurun_out/r2ac_ncu.log 2>&1; tail -2 gpurun_out/r2ac_ncu.log
