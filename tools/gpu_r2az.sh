# round 2, last call: the library was rebuilt from the same sources after a reverted experiment (rebuilds are not bit-reproducible), so
# the evidence tied to the binary's hash is taken again: executed counts, full GPU suite, default bench
set -x
mkdir -p gpurun_out /tmp/rep
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum
PROF_TAPE_GB=24 PROF_COUNTS=gpurun_out/r2az_prof_counts.json timeout 900 ncu --metrics $M --clock-control none -k regex:"k_ros_fwd_cell_tm|k_expand_tape|k_ros_adj_pair" -o /tmp/rep/exec python tools/prof_case.py 2368 12 > gpurun_out/r2az_ncu_exec.log 2>&1; tail -1 gpurun_out/r2az_ncu_exec.log
python tools/ncu_executed.py /tmp/rep/exec.ncu-rep gpurun_out/r2az_prof_counts.json fatode_b200/lib/libfatode_b200.so "tools/gpu_r2az.sh: prof_case.py 2368 12" > gpurun_out/executed_counts.json; head -c 300 gpurun_out/executed_counts.json
cp gpurun_out/executed_counts.json profiles/executed_counts.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2az_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2az_pytest_gpu.log; tail -3 gpurun_out/r2az_pytest_gpu.log
( time timeout 1500 python bench.py > gpurun_out/r2az_bench_default.json 2> gpurun_out/r2az_bench_default.err ) 2> gpurun_out/r2az_time_default.txt; tail -c 300 gpurun_out/r2az_bench_default.json; tail -3 gpurun_out/r2az_time_default.txt
du -sh gpurun_out
