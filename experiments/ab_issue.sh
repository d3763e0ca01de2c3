#!/bin/bash
# A/B within ONE box session (boxes differ by +-5 % under the power cap): single-lane issue
# (default build) vs converged-warp issue (-DDTX_CONVERGED_ISSUE), kblk 2 and 4, interleaved.
mkdir -p gpurun_out
: > gpurun_out/ab_issue.txt
for rep in 1 2; do
for k in 2 4; do
for v in default conv; do
  if [ $v = conv ]; then export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_conv.so; else unset DETEX_B200_LIB; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --kblk $k --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep variant $v kblk $k value %.4g k1_ms %.1f sm_mhz %s' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz']))
" | tee -a gpurun_out/ab_issue.txt
done; done; done
