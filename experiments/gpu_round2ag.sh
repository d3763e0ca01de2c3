#!/bin/bash
# round 2, call AG: last code of the round: whole GPU suite + smoke with the default library, CCX timing, racecheck of
# the tool-sync build of the ring kernel (-DDTX_RING_TOOL_SYNC=1) and of the default build
mkdir -p gpurun_out
( python -m pytest tests -m gpu -q ) > gpurun_out/r2ag_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ag_pytest.log; tail -3 gpurun_out/r2ag_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --sections ccx --no-cpu --no-alt --chunks 24 2> gpurun_out/r2ag_ccx.err | tail -1 > gpurun_out/r2ag_ccx.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ag_ccx.json').read())['ccx']
print('ccx resident %.1f ms  e2e %.1f ms  k1 %.1f ms other %.1f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1']))
P
DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_toolsync.so timeout 140 compute-sanitizer --tool racecheck --print-limit 20 python experiments/sanitize_ccx.py > gpurun_out/r2ag_racecheck_toolsync.log 2>&1
grep -E "RACECHECK SUMMARY|sanitize target ok" gpurun_out/r2ag_racecheck_toolsync.log | tail -2
timeout 140 compute-sanitizer --tool racecheck --print-limit 20 python experiments/sanitize_ccx.py > gpurun_out/r2ag_racecheck_default.log 2>&1
grep -E "RACECHECK SUMMARY|sanitize target ok" gpurun_out/r2ag_racecheck_default.log | tail -2
timeout 140 compute-sanitizer --tool memcheck --print-limit 20 python experiments/sanitize_ccx.py > gpurun_out/r2ag_memcheck_default.log 2>&1
grep -E "ERROR SUMMARY|sanitize target ok" gpurun_out/r2ag_memcheck_default.log | tail -2
