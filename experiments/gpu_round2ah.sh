#!/bin/bash
# round 2, call AH (4 GPUs): configs[2] at 4 GPUs with the final defaults (16 GiB series buffer)
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
( time $TR bench.py --gpus $N --sections ccx --ccx-events 4096 --chunks 24 --no-alt --no-cpu ) > gpurun_out/r2ah_ccx4096_n$N.json 2> gpurun_out/r2ah_ccx4096_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r2ah_ccx4096_n$N.err
python - $N <<'P'
import json,sys
n=sys.argv[1]
d=json.loads(open('gpurun_out/r2ah_ccx4096_n%s.json'%n).read().strip().splitlines()[-1]); c=d['ccx']; e=c['e2e']
print('4096 @%s GPUs: resident %.1f ms  e2e(host buffer) %.1f ms  nccl all %.1f  nccl root %.1f  k1 %.1f'%(n,c['ms_per_step'],e['ms_per_step'],e['ms_per_step_nccl_gather_every_rank'],e['ms_per_step_nccl_gather_rank0_only'],c['roofline']['k1_ms_per_call']))
P
