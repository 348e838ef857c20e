# round 2, call ar (2 GPUs): bench under torchrun with the HEAD library on the in-library NCCL communicator, weak (2^17 cells per GPU, the
# driver's size) and strong (2^18 cells split), plus the C host
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_c_host.py -m gpu -x -q -s > gpurun_out/r2ar_c_host.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ar_c_host.log; tail -4 gpurun_out/r2ar_c_host.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cells 131072 --steps 3 --warmup 3 > gpurun_out/r2ar_bench_2gpu_131072.json 2> gpurun_out/r2ar_bench_2gpu.err; tail -c 500 gpurun_out/r2ar_bench_2gpu_131072.json; tail -3 gpurun_out/r2ar_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --scaling strong --cells 262144 --steps 2 --warmup 3 --no-e2e > gpurun_out/r2ar_bench_2gpu_strong_262144.json 2> gpurun_out/r2ar_bench_2gpu_strong.err; tail -c 300 gpurun_out/r2ar_bench_2gpu_strong_262144.json
