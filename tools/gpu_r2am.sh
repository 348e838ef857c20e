# round 2, final evidence call: full GPU suite, default bench + reference arm, ncu launch lists, executed counts and --set full captures
# of the three headline kernels from the HEAD binary, --set full of the new kernels
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2am_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2am_pytest_gpu.log; tail -3 gpurun_out/r2am_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2am_bench_default.json 2> gpurun_out/r2am_bench_default.err ) 2> gpurun_out/r2am_time_default.txt; tail -c 300 gpurun_out/r2am_bench_default.json; tail -3 gpurun_out/r2am_time_default.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2am_bench_reference.json 2> gpurun_out/r2am_bench_reference.err; tail -c 300 gpurun_out/r2am_bench_reference.json
timeout 600 python tools/all_entry_points_case.py > gpurun_out/r2am_all_entry_points.log 2>&1; tail -8 gpurun_out/r2am_all_entry_points.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2am_bench_launches.csv python bench.py --cells 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-side > gpurun_out/r2am_ncu_bench.log 2>&1; tail -1 gpurun_out/r2am_ncu_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2am_all_entry_points_launches.csv python tools/all_entry_points_case.py > gpurun_out/r2am_ncu_all.log 2>&1; tail -1 gpurun_out/r2am_ncu_all.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum
PROF_TAPE_GB=24 PROF_COUNTS=gpurun_out/r2am_prof_counts.json timeout 900 ncu --metrics $M --clock-control none -k regex:"k_ros_fwd_cell_tm|k_expand_tape|k_ros_adj_pair" -o gpurun_out/prof_r2am_exec python tools/prof_case.py 2368 12 > gpurun_out/r2am_ncu_exec.log 2>&1; tail -1 gpurun_out/r2am_ncu_exec.log
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_adj_pair" -c 1 -o gpurun_out/prof_r2am_adj_pair python tools/prof_case.py 2368 12 > gpurun_out/r2am_ncu_adj.log 2>&1; tail -1 gpurun_out/r2am_ncu_adj.log
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_fwd_cell_tm" -c 1 -o gpurun_out/prof_r2am_fwd_tm python tools/prof_case.py 8584 12 > gpurun_out/r2am_ncu_fwd.log 2>&1; tail -1 gpurun_out/r2am_ncu_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_erk_adj_swe" -c 1 -o gpurun_out/prof_r2am_erk_adj python tools/erk_adj_case.py 296 adj > gpurun_out/r2am_ncu_erk_adj.log 2>&1; tail -1 gpurun_out/r2am_ncu_erk_adj.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sdirk_tlm_lock" -c 1 -o gpurun_out/prof_r2am_sdirk_lock python tools/sdirk_lock_case.py 296 direct 6 > gpurun_out/r2am_ncu_lock.log 2>&1; tail -1 gpurun_out/r2am_ncu_lock.log
ls -la gpurun_out/*.ncu-rep | tail -8
