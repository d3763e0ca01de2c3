#!/bin/bash
# epilogue with smem transposition (coalesced 512-byte DS rows, slot sums in the read-out instead of
# shuffles): GPU tests, then same-box A/B against the previous epilogue for both engines
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
: > gpurun_out/ab_epi2.txt
for rep in 1 2; do
for e in tcgen05 tcgen05_x8; do
for v in prev new; do
  if [ $v = prev ]; then export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_prev.so; else unset DETEX_B200_LIB; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --engine $e --no-cpu --no-alt 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep engine $e epilogue $v value %.4g k1_ms %.1f sm_mhz %s parity %.3g' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz'], d['parity_check']['max_abs_err_vs_fp64']))
" | tee -a gpurun_out/ab_epi2.txt
done; done; done
unset DETEX_B200_LIB
echo "== CCX check"; timeout 600 python experiments/ccx_bench.py 2>&1 | tail -4
