#!/bin/bash
# round 2, call AA: length of the first two TMEM accumulations of a detection work item (the MMA warp's slack while the
# drain warps are still in the previous item's epilogue): factor 2 (default) / 3 / 4 / 5 x kblk, same-box A/B
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
for rep in 1 2; do
  for f in 2 3 4 5; do
    lib=detex_b200/_C/libdetex_b200_f$f.so; [ $f = 2 ] && lib=detex_b200/_C/libdetex_b200.so
    DETEX_B200_LIB=$PWD/$lib $B > gpurun_out/r2aa_f${f}_$rep.json 2>> gpurun_out/r2aa_err.log
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2aa_f*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'parity %.3g'%d['parity_check']['max_abs_err_vs_fp64'], 'clk', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2aa_err.log
# the long-event CCX fallback test added after the last full run
python -m pytest tests/test_gpu_ccx.py -m gpu -q -k "long_events or pack_rows" 2>&1 | tail -2
