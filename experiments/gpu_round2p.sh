#!/bin/bash
# round 2, call P: fused epilogue with the per-value path out of line (code size 31 k -> 14 k SASS), same-box A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -q > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log; tail -4 gpurun_out/r2p_pytest.log
B="python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
for rep in 1 2; do
  $B > gpurun_out/r2p_unfused_$rep.json 2>> gpurun_out/r2p_err.log
  $B --fused > gpurun_out/r2p_fused_$rep.json 2>> gpurun_out/r2p_err.log
done
$B --fused --batch 96 > gpurun_out/r2p_fused_b96.json 2>> gpurun_out/r2p_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'share %.4f'%d['roofline']['k1_share_of_step'], 'cands', d['candidates_per_step'], 'hist', d['hist_total'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2p_err.log
