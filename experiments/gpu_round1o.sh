#!/bin/bash
# same-box A/B, all with long first accumulations + streaming DS stores:
#  vF  = scalar mu loads, atomics inline (commit d1e4014's epilogue)
#  vG  = vector mu loads (phase-major mu tile), atomics inline
#  new = vector mu loads, rare atomics path out of line (default build)
# then ncu --set full of K1 (default build) for the DRAM traffic after the streaming stores
mkdir -p gpurun_out
: > gpurun_out/ab_var2.txt
for e in tcgen05 tcgen05_x8; do
for v in vF vG new; do
  if [ $v = new ]; then unset DETEX_B200_LIB; else export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_$v.so; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --engine $e --no-cpu --no-alt 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('engine $e variant $v value %.4g k1_ms %.1f sm_mhz %s parity %.3g' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz'], d['parity_check']['max_abs_err_vs_fp64']))
" | tee -a gpurun_out/ab_var2.txt
done; done
unset DETEX_B200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_tcgen05_r1o -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/ncu_k1_r1o.log 2>&1
tail -1 gpurun_out/ncu_k1_r1o.log | cut -c1-150
