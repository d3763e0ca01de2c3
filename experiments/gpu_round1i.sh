#!/bin/bash
# adaptive 8-bit cross-term engine: GPU tests, model calibration, A/B speed in one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
timeout 600 python experiments/x8_check.py > gpurun_out/x8_check.txt 2>&1; tail -12 gpurun_out/x8_check.txt
: > gpurun_out/ab_x8.txt
for e in tcgen05 tcgen05_auto tcgen05_x8; do
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --engine $e --no-cpu 2>gpurun_out/ab_$e.err | tee gpurun_out/ab_$e.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('engine $e value %.4g e2e %.4g k1_ms %.1f sm_mhz %s frac %.3f parity %s x8 %s' % (d['value'], d['e2e']['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['parity_check'], d.get('x8_chunk_fraction')))
" | tee -a gpurun_out/ab_x8.txt
done
