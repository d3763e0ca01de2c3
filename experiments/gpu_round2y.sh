#!/bin/bash
# round 2, call Y (N GPUs): shared-host CCX result path (dtx_ccx_pack_rows into a /dev/shm matrix every rank has
# page-locked) against the NCCL gather; CCX tests with the trimmed scan
mkdir -p gpurun_out
N=${1:-2}
( CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py -m gpu -q -x -k "ccx" ) > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -3 gpurun_out/r2y_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for ev in 4096 16384; do
  ( time $TR bench.py --gpus $N --sections ccx --ccx-events $ev --chunks 24 --no-alt --no-cpu ) > gpurun_out/r2y_ccx${ev}_n$N.json 2> gpurun_out/r2y_ccx${ev}_n$N.err
  echo "rc=$?" >> gpurun_out/r2y_ccx${ev}_n$N.err
  tail -4 gpurun_out/r2y_ccx${ev}_n$N.err
  python - $N $ev <<'P'
import json,sys
n,ev=sys.argv[1:3]
try:
    d=json.loads(open('gpurun_out/r2y_ccx%s_n%s.json'%(ev,n)).read().strip().splitlines()[-1]); c=d['ccx']; e=c['e2e']
    print(ev,'resident %.1f ms  e2e(host buffer) %.1f ms  nccl all %.1f  nccl root %.1f  k1 %.1f'%(c['ms_per_step'],e['ms_per_step'],e['ms_per_step_nccl_gather_every_rank'],e['ms_per_step_nccl_gather_rank0_only'],c['roofline']['k1_ms_per_call']))
except Exception as ex: print(ev,'failed',ex)
P
done
ls /dev/shm | head -3
