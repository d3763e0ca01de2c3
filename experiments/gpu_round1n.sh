#!/bin/bash
# same-box A/B of epilogue / accumulation variants (one repetition, both engines):
#  prev = commit d1e4014 (transposed epilogue, scalar mu loads, atomics inline)
#  vB   = vector mu loads + rare atomics path out of line
#  vC   = vB + streaming DS stores        vD = vB + long first accumulations
#  new  = vB + both (the default build)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
: > gpurun_out/ab_var.txt
for e in tcgen05 tcgen05_x8; do
for v in prev vB vC vD new; do
  if [ $v = new ]; then unset DETEX_B200_LIB; else export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_$v.so; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --engine $e --no-cpu --no-alt 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('engine $e variant $v value %.4g k1_ms %.1f sm_mhz %s parity %.3g' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz'], d['parity_check']['max_abs_err_vs_fp64']))
" | tee -a gpurun_out/ab_var.txt
done; done
unset DETEX_B200_LIB
