# round 2: 4 GPUs, weak-scaling bench line with the shipped library (completes the 1 / 2 / 4 / 8 series)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --cells 65536 --steps 2 --warmup 3 > gpurun_out/r2ba_bench_4gpu_65536.json 2> gpurun_out/r2ba_bench_4gpu.err; tail -c 300 gpurun_out/r2ba_bench_4gpu_65536.json
