#!/bin/bash
# round 2, call Q: CTA pairs (cta_group::2) for K1 -- correctness under a timeout, then same-box A/B
mkdir -p gpurun_out
DTX_K1_CG2=1 timeout 180 python -m pytest tests/test_gpu_detect.py -m gpu -q -x -k "golden or rank16 or histogram" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -8 gpurun_out/r2q_pytest.log
if grep -q "rc=0" gpurun_out/r2q_pytest.log; then
  DTX_K1_CG2=1 timeout 300 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fused.py tests/test_gpu_x8.py tests/test_gpu_scale.py -m gpu -q -k "not ccx" > gpurun_out/r2q_pytest2.log 2>&1
  echo "pytest2 rc=$?" >> gpurun_out/r2q_pytest2.log; tail -5 gpurun_out/r2q_pytest2.log
  B="timeout 300 python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
  for rep in 1 2; do
    $B > gpurun_out/r2q_cg1_$rep.json 2>> gpurun_out/r2q_err.log
    DTX_K1_CG2=1 $B > gpurun_out/r2q_cg2_$rep.json 2>> gpurun_out/r2q_err.log
  done
  python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_cg*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'parity %.3g'%d['parity_check']['max_abs_err_vs_fp64'], 'cands', d['candidates_per_step'], 'hist', d['hist_total'], 'clk', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e)
PY
  tail -5 gpurun_out/r2q_err.log
fi
nvidia-smi --query-gpu=name,utilization.gpu --format=csv | tail -1
