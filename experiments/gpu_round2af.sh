#!/bin/bash
# round 2, call AF: ring kernel with per-lane arrives / atomic gate: CCX tests, racecheck + memcheck, timing
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py -m gpu -q -x -k "ccx or pack or long_events" ) > gpurun_out/r2af_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log; tail -3 gpurun_out/r2af_pytest.log
for tool in racecheck memcheck; do
  timeout 140 compute-sanitizer --tool $tool --print-limit 20 python experiments/sanitize_ccx.py > gpurun_out/r2af_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2af_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|sanitize target ok" gpurun_out/r2af_$tool.log | tail -3
done
python bench.py --sections ccx --no-cpu --no-alt --chunks 24 2> gpurun_out/r2af_ccx.err | tail -1 > gpurun_out/r2af_ccx.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2af_ccx.json').read())['ccx']
print('ccx resident %.1f ms  e2e %.1f ms  k1 %.1f ms other %.1f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1']))
P
