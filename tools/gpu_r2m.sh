# round 2, call m: ncu --set full of the two-warps-per-cell backward kernel (2368 cells = one round of 148 SMs x 4 pairs, 12 h window)
set -x
mkdir -p gpurun_out
PROF_TAPE_GB=24 timeout 300 python tools/prof_case.py 2368 12 > gpurun_out/r2m_prof_plain.log 2>&1; tail -1 gpurun_out/r2m_prof_plain.log
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_adj_pair" -c 1 -o gpurun_out/prof_r2m_adj_pair python tools/prof_case.py 2368 12 > gpurun_out/r2m_ncu.log 2>&1; tail -2 gpurun_out/r2m_ncu.log
