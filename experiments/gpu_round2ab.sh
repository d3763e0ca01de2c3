#!/bin/bash
# round 2, call AB: cfg1 end to end in three overlapped batches; CCX signal-batch size (series buffer 4 / 8 / 16 GiB)
mkdir -p gpurun_out
python bench.py --sections cfg1 --chunks 24 --no-cpu --no-alt 2> gpurun_out/r2ab_cfg1.err | tail -1 > gpurun_out/r2ab_cfg1.json
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2ab_cfg1.json').read())['cfg1']
    print('cfg1 resident %.4g (%.2f ms)  e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
except Exception as e: print('cfg1 failed',e); print(open('gpurun_out/r2ab_cfg1.err').read()[-600:])
P
for g in 4 8 16; do
  python bench.py --sections ccx --no-cpu --no-alt --chunks 24 --ccx-ds-gib $g 2> gpurun_out/r2ab_ccx$g.err | tail -1 > gpurun_out/r2ab_ccx$g.json
  python - $g <<'P'
import json,sys
g=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2ab_ccx%s.json'%g).read())['ccx']
    print('ds %s GiB: resident %.1f ms  e2e %.1f ms  k1 %.1f ms (%d launches) other %.1f'%(g,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['roofline']['launches_per_call'],d['gpu_ms_other_than_k1']))
except Exception as e: print(g,'failed',e); print(open('gpurun_out/r2ab_ccx%s.err'%g).read()[-600:])
P
done
