#!/bin/bash
# round 2, call X (8 GPUs): every bench section with the ring re-scoring + K1 DUAL (reduced station: 96 chunks per GPU),
# then configs[2] at 16384 events
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
( time $TR bench.py --gpus $N --steps 2 --warmup 3 --chunks 96 --no-alt ) > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err
echo "rc=$?" >> gpurun_out/r2x_bench_n$N.err
tail -5 gpurun_out/r2x_bench_n$N.err
( time $TR bench.py --gpus $N --sections ccx --ccx-events 16384 --chunks 24 --no-alt ) > gpurun_out/r2x_ccx16k_n$N.json 2> gpurun_out/r2x_ccx16k_n$N.err
echo "rc=$?" >> gpurun_out/r2x_ccx16k_n$N.err
tail -5 gpurun_out/r2x_ccx16k_n$N.err
python - $N <<'P'
import json,sys
n=sys.argv[1]
for f in ('bench','ccx16k'):
    try:
        d=json.loads(open('gpurun_out/r2x_%s_n%s.json'%(f,n)).read().strip().splitlines()[-1])
        c=d['ccx']
        print(f,'ccx resident %.1f ms e2e %.1f ms (root only %.1f) k1 %.1f'%(c['ms_per_step'],c['e2e']['ms_per_step'],c['e2e'].get('ms_per_step_result_on_rank0_only',-1),c['roofline']['k1_ms_per_call']))
        if 'value' in d: print(' main %.4g ts/s e2e %.4g; fas %.4g'%(d['value'],d['e2e']['value'],d.get('fas',{}).get('value',0)))
    except Exception as e: print(f,'failed',e)
P
