#!/bin/bash
# ncu --set full of K1 after the epilogue change, both precision modes (48-chunk launch)
mkdir -p gpurun_out
for e in tcgen05 tcgen05_x8; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_${e}_r1m -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt --engine $e > gpurun_out/ncu_k1_${e}.log 2>&1
tail -1 gpurun_out/ncu_k1_${e}.log | cut -c1-150
done
