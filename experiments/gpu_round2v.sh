#!/bin/bash
# round 2, call V: 8-stage A ring in hi-only mode; CCX tests; configs[2] at 4096 and 16384 events on one GPU
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py tests/test_gpu_case1_workflow.py tests/test_dropin.py -m gpu -q -x ) > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log; tail -4 gpurun_out/r2v_pytest.log
run() {
  tag=$1; ev=$2; shift; shift
  env "$@" python bench.py --sections ccx --no-cpu --no-alt --chunks 24 --ccx-events $ev 2> gpurun_out/r2v_$tag.err | tail -1 > gpurun_out/r2v_$tag.json
  python - "$tag" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2v_%s.json'%t).read())['ccx']
    print('%-14s resident %.1f ms  e2e %.1f ms  k1 %.1f ms  other %.1f ms  frac %.3f'%(t,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1'],d['roofline']['frac']))
except Exception as e:
    print(t,'failed',e); print(open('gpurun_out/r2v_%s.err'%t).read()[-800:])
P
}
run n4096 4096 DTX_X=0
run n4096b 4096 DTX_X=0
run n16384 16384 DTX_X=0
ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 20 -c 1 -o gpurun_out/r2v_k1dual python bench.py --sections ccx --no-cpu --no-alt --chunks 24 > gpurun_out/r2v_ncu.log 2>&1
