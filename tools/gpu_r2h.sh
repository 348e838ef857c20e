# round 2, call h (2 GPUs): plain C host with the in-library NCCL all-reduce; bench under torchrun using the same communicator
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_c_host.py -m gpu -x -q -s > gpurun_out/r2h_c_host.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_c_host.log; tail -6 gpurun_out/r2h_c_host.log
tests/c/abi_two_ranks 2 64 6.0 > gpurun_out/r2h_c_host_run.log 2>&1; cat gpurun_out/r2h_c_host_run.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cells 32768 --steps 2 --warmup 1 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err; tail -c 600 gpurun_out/r2h_bench_2gpu.json; tail -3 gpurun_out/r2h_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --scaling strong --cells 65536 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2h_bench_2gpu_strong.json 2> gpurun_out/r2h_bench_2gpu_strong.err; tail -c 300 gpurun_out/r2h_bench_2gpu_strong.json
