#!/bin/bash
# round 2, call W: what starves the MMA warp of K1 DUAL (64 % of its samples wait for an A stage)?  Item order sweep:
# with group = 128 all CTAs stream the same 786 KB block at once; smaller chunk groups spread them over more blocks
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python bench.py --sections ccx --no-cpu --no-alt --chunks 24 2> gpurun_out/r2w_$tag.err | tail -1 > gpurun_out/r2w_$tag.json
  python - "$tag" <<'P'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2w_%s.json'%t).read())['ccx']
    print('%-14s resident %.1f ms  e2e %.1f ms  k1 %.1f ms  other %.1f ms  frac %.3f'%(t,d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['k1_ms_per_call'],d['gpu_ms_other_than_k1'],d['roofline']['frac']))
except Exception as e:
    print(t,'failed',e); print(open('gpurun_out/r2w_%s.err'%t).read()[-800:])
P
}
run g128 DTX_X=0
run g64 DTX_K1_GROUP=64
run g32 DTX_K1_GROUP=32
run g16 DTX_K1_GROUP=16
run g8 DTX_K1_GROUP=8
run g4 DTX_K1_GROUP=4
run g2 DTX_K1_GROUP=2
