#!/bin/bash
# round 2, call J: evidence for profiles/ -- launch list of the bench command, one full ncu capture of K1
# (collector-reuse build), kblk A/B on the same box
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
for k in 2 3 4 2; do
  $B --kblk $k > gpurun_out/r2j_kblk$k.json 2>> gpurun_out/r2j_err.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k0_|k1_|k3_|lta|basis|zero_energy|sum_pieces" -c 300 --csv --log-file gpurun_out/r2j_ncu_launches.csv $B > gpurun_out/r2j_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_kernel" -s 2 -c 1 -o gpurun_out/r2j_k1_full python bench.py --steps 1 --warmup 1 --chunks 48 --no-cpu --no-alt --sections main > gpurun_out/r2j_ncu2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_kblk*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'parity', d['parity_check']['max_abs_err_vs_fp64'], 'clk', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e)
PY
ls -la gpurun_out/r2j_k1_full.ncu-rep; tail -3 gpurun_out/r2j_err.log
