#!/bin/bash
# round 2, call K: fused epilogue (tests, A/B against the unfused path on the same box), full ncu capture of K1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_detect.py -m gpu -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -15 gpurun_out/r2k_pytest.log
B="python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
$B > gpurun_out/r2k_unfused.json 2>> gpurun_out/r2k_err.log
$B --fused > gpurun_out/r2k_fused.json 2>> gpurun_out/r2k_err.log
$B --fused --batch 96 > gpurun_out/r2k_fused_b96.json 2>> gpurun_out/r2k_err.log
$B > gpurun_out/r2k_unfused2.json 2>> gpurun_out/r2k_err.log
ncu --set full --clock-control none --import-source on -k regex:"k1_kernel" -s 1 -c 1 -o gpurun_out/r2k_k1_full python bench.py --steps 1 --warmup 1 --chunks 48 --no-cpu --no-alt --sections main > gpurun_out/r2k_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_kernel" -s 1 -c 1 -o gpurun_out/r2k_k1_fused_full python bench.py --steps 1 --warmup 1 --chunks 48 --no-cpu --no-alt --sections main --fused > gpurun_out/r2k_ncu2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'share %.4f'%d['roofline']['k1_share_of_step'], 'cands', d['candidates_per_step'], 'hist', d['hist_total'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2k_err.log
