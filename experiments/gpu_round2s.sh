#!/bin/bash
# round 2, call S: ring re-scoring (producer / consumer ring, dynamic tasks, one-pass multi-candidate) + single-pass
# scan: CCX parity tests, drop-in tests against the unmodified reference, A/B of the knobs on configs[2]
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py tests/test_dropin.py tests/test_gpu_case1_workflow.py -m gpu -q -x ) > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log; tail -5 gpurun_out/r2s_pytest.log
run() {
  tag=$1; shift
  env "$@" python bench.py --sections ccx --no-cpu --no-alt --chunks 24 2> gpurun_out/r2s_$tag.err | tail -1 > gpurun_out/r2s_$tag.json
  python - "$tag" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2s_%s.json'%t).read())['ccx']
    print('%-14s resident %.1f ms  e2e %.1f ms  k1 %.1f ms  other %.1f ms'%(t,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1']))
except Exception as e:
    print(t,'failed',e); print(open('gpurun_out/r2s_%s.err'%t).read()[-800:])
P
}
run tiled DTX_CCX_POST_TILED=1
run ring DTX_X=0
run ring_w16 DTX_CCX_RING_WARPS=16
run ring_w9 DTX_CCX_RING_WARPS=9
run ring_nb2 DTX_CCX_RING_NB=2 DTX_CCX_RING_TC=7
run ring_nb4 DTX_CCX_RING_NB=4 DTX_CCX_RING_TC=5
run ring_nb2w16 DTX_CCX_RING_NB=2 DTX_CCX_RING_TC=7 DTX_CCX_RING_WARPS=16
run ring_rg128 DTX_CCX_RING_RG=128
run ring_rg32 DTX_CCX_RING_RG=32
