#!/bin/bash
# 2 GPUs: CCX row-block sharding + NCCL gather (configs[2]), detection bench at N=2 with the final kernels
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   experiments/ccx_multi_gpu.py 4096 2>&1 | grep '^{' | tee gpurun_out/ccx_n2.json | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu --no-alt 2>gpurun_out/bench_n2.err | grep '^{' | tee gpurun_out/bench_n2.json | cut -c1-400
tail -2 gpurun_out/bench_n2.err
