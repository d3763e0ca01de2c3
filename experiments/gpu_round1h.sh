#!/bin/bash
# x8 engine: correctness probe, then A/B speed in one box
mkdir -p gpurun_out
timeout 600 python experiments/x8_check.py > gpurun_out/x8_check.txt 2>&1; tail -12 gpurun_out/x8_check.txt
: > gpurun_out/ab_x8.txt
for rep in 1 2; do
for e in tcgen05 tcgen05_x8; do
for k in 2 4; do
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --kblk $k --engine $e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep engine $e kblk $k value %.4g k1_ms %.1f sm_mhz %s' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz']))
" | tee -a gpurun_out/ab_x8.txt
done; done; done
