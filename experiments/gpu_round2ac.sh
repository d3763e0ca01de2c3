#!/bin/bash
# round 2, call AC: 16 GiB series buffer as the default (CCX tests, 4096 / 16384 events), cfg1 end to end in three
# overlapped batches (needs the main section's host copy)
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py tests/test_dropin.py -m gpu -q -x ) > gpurun_out/r2ac_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log; tail -3 gpurun_out/r2ac_pytest.log
python bench.py --sections main,cfg1 --chunks 48 --steps 1 --warmup 1 --no-cpu --no-alt 2> gpurun_out/r2ac_cfg1.err | tail -1 > gpurun_out/r2ac_cfg1.json
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2ac_cfg1.json').read())['cfg1']
    print('cfg1 resident %.4g (%.2f ms)  e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
except Exception as e: print('cfg1 failed',e); print(open('gpurun_out/r2ac_cfg1.err').read()[-600:])
P
for ev in 4096 16384; do
  python bench.py --sections ccx --no-cpu --no-alt --chunks 24 --ccx-events $ev 2> gpurun_out/r2ac_ccx$ev.err | tail -1 > gpurun_out/r2ac_ccx$ev.json
  python - $ev <<'P'
import json,sys
g=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2ac_ccx%s.json'%g).read())['ccx']
    print('%s events: resident %.1f ms  e2e %.1f ms  k1 %.1f ms (%d launches) other %.1f frac %.3f'%(g,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['roofline']['launches_per_call'],d['gpu_ms_other_than_k1'],d['roofline']['frac']))
except Exception as e: print(g,'failed',e); print(open('gpurun_out/r2ac_ccx%s.err'%g).read()[-600:])
P
done
