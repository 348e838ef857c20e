# round 2, call aw (8 GPUs): weak-scaling bench line with the HEAD library on the in-library NCCL communicator
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --cells 65536 --steps 2 --warmup 3 > gpurun_out/r2aw_bench_8gpu_65536.json 2> gpurun_out/r2aw_bench_8gpu.err; tail -c 400 gpurun_out/r2aw_bench_8gpu_65536.json; tail -2 gpurun_out/r2aw_bench_8gpu.err
