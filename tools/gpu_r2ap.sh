# round 2, final evidence (2/3): ncu captures of the three headline kernels from the HEAD binary, summarised ON THE BOX (the
# reports are too large to travel): executed counts, --set full summaries and per-function / per-line tables
set -x
mkdir -p gpurun_out /tmp/rep
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum
PROF_TAPE_GB=24 PROF_COUNTS=gpurun_out/r2ap_prof_counts.json timeout 900 ncu --metrics $M --clock-control none -k regex:"k_ros_fwd_cell_tm|k_expand_tape|k_ros_adj_pair" -o /tmp/rep/exec python tools/prof_case.py 2368 12 > gpurun_out/r2ap_ncu_exec.log 2>&1; tail -1 gpurun_out/r2ap_ncu_exec.log
python tools/ncu_executed.py /tmp/rep/exec.ncu-rep gpurun_out/r2ap_prof_counts.json fatode_b200/lib/libfatode_b200.so "tools/gpu_r2ap.sh: prof_case.py 2368 12" > gpurun_out/executed_counts.json; head -c 600 gpurun_out/executed_counts.json
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_adj_pair" -c 1 -o /tmp/rep/adj_pair python tools/prof_case.py 2368 12 > gpurun_out/r2ap_ncu_adj.log 2>&1
python tools/ncu_summary.py /tmp/rep/adj_pair.ncu-rep > gpurun_out/r2ap_adj_pair_summary.txt 2>&1; python tools/ncu_roles.py /tmp/rep/adj_pair.ncu-rep >> gpurun_out/r2ap_adj_pair_summary.txt 2>&1; tail -5 gpurun_out/r2ap_adj_pair_summary.txt
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ros_fwd_cell_tm" -c 1 -o /tmp/rep/fwd_tm python tools/prof_case.py 8584 12 > gpurun_out/r2ap_ncu_fwd.log 2>&1
python tools/ncu_summary.py /tmp/rep/fwd_tm.ncu-rep > gpurun_out/r2ap_fwd_tm_summary.txt 2>&1; python tools/ncu_funcs.py /tmp/rep/fwd_tm.ncu-rep k_ros_fwd_cell_tm fatode_b200/csrc/gen/cbm4_cell_gen.h >> gpurun_out/r2ap_fwd_tm_summary.txt 2>&1; tail -5 gpurun_out/r2ap_fwd_tm_summary.txt
PROF_TAPE_GB=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_expand_tape" -c 1 -o /tmp/rep/expand python tools/prof_case.py 2368 12 > gpurun_out/r2ap_ncu_expand.log 2>&1
python tools/ncu_summary.py /tmp/rep/expand.ncu-rep > gpurun_out/r2ap_expand_summary.txt 2>&1
du -sh gpurun_out
