# round 2, call a: compute-sanitizer over every kernel family (HEAD binary) + fresh ncu captures of the three headline kernels
set -x
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_smoke_case.py 1.5 8 > gpurun_out/r2a_san_${tool}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_san_${tool}_smoke.log; tail -4 gpurun_out/r2a_san_${tool}_smoke.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/all_entry_points_case.py > gpurun_out/r2a_san_${tool}_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_san_${tool}_all.log; tail -4 gpurun_out/r2a_san_${tool}_all.log
done
PROF_TAPE_GB=24 timeout 300 python tools/prof_case.py 2368 12 > gpurun_out/r2a_prof_plain.log 2>&1; tail -1 gpurun_out/r2a_prof_plain.log
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_fwd_cell|k_expand_tape|k_ros_adj_cell" -c 3 -o gpurun_out/prof_r2a_cellpath python tools/prof_case.py 2368 12 > gpurun_out/r2a_ncu.log 2>&1; tail -2 gpurun_out/r2a_ncu.log
