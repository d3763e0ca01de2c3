#!/bin/bash
# round 2, call U: K1 DUAL (CCX screening series on pairs of signals sharing the basis-image stream): whole GPU suite,
# A/B on configs[2], ncu of the dual kernel
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -5 gpurun_out/r2u_pytest.log
run() {
  tag=$1; shift
  env "$@" python bench.py --sections ccx --no-cpu --no-alt --chunks 24 2> gpurun_out/r2u_$tag.err | tail -1 > gpurun_out/r2u_$tag.json
  python - "$tag" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2u_%s.json'%t).read())['ccx']
    print('%-14s resident %.1f ms  e2e %.1f ms  k1 %.1f ms  other %.1f ms  frac %.3f'%(t,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1'],d['roofline']['frac']))
except Exception as e:
    print(t,'failed',e); print(open('gpurun_out/r2u_%s.err'%t).read()[-800:])
P
}
run single DTX_K1_NODUAL=1
run dual DTX_X=0
run dual2 DTX_X=0
ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 20 -c 1 -o gpurun_out/r2u_k1dual python bench.py --sections ccx --no-cpu --no-alt --chunks 24 > gpurun_out/r2u_ncu.log 2>&1
DTX_K1_NODUAL=1 ncu --set full --clock-control none -k regex:k1_kernel -s 20 -c 1 -o gpurun_out/r2u_k1single python bench.py --sections ccx --no-cpu --no-alt --chunks 24 > gpurun_out/r2u_ncu1.log 2>&1
ls -la gpurun_out/r2u_*rep
