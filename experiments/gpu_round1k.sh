#!/bin/bash
# (1) GPU tests after the K3 histogram change; (2) same-box A/B of the issue style (single lane vs
# converged warp with an elected lane) for the 3-MMA and the 8-bit cross-term engines
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
: > gpurun_out/ab_issue2.txt
for rep in 1 2; do
for e in tcgen05 tcgen05_x8; do
for v in default conv; do
  if [ $v = conv ]; then export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_conv.so; else unset DETEX_B200_LIB; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --engine $e --no-cpu --no-alt 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep engine $e variant $v value %.4g k1_ms %.1f sm_mhz %s parity %.3g' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz'], d['parity_check']['max_abs_err_vs_fp64']))
" | tee -a gpurun_out/ab_issue2.txt
done; done; done
unset DETEX_B200_LIB
echo "== ncu launch list"
KREG='regex:k0_|k1_kernel|k3_|lta_kernel|basis_image'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
