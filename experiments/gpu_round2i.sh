#!/bin/bash
# round 2, call I (2 GPUs): every bench section under torchrun (NCCL): sharded CCX, sharded FAS, per-rank stats
mkdir -p gpurun_out
N=${1:-2}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 --chunks 96 ) > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err
echo "rc=$?" >> gpurun_out/r2i_bench_n$N.err
tail -8 gpurun_out/r2i_bench_n$N.err
