#!/bin/bash
# round 2, call AD (2 GPUs): the whole GPU suite on GPU 0 with the final code, smoke, then every bench section under
# torchrun at 2 GPUs (reduced station) -- the bench asserts that the resident / host paths and the shared-host / NCCL
# CCX results agree
mkdir -p gpurun_out
( CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -q ) > gpurun_out/r2ad_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.log; tail -3 gpurun_out/r2ad_pytest.log
CUDA_VISIBLE_DEVICES=0 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
( time $TR bench.py --gpus 2 --steps 2 --warmup 3 --chunks 48 --no-alt ) > gpurun_out/r2ad_bench_n2.json 2> gpurun_out/r2ad_bench_n2.err
echo "bench rc=$?"; tail -4 gpurun_out/r2ad_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2ad_bench_n2.json').read().strip().splitlines()[-1])
c=d['ccx']
print('main %.4g e2e %.4g | cfg1 %.4g e2e %.4g | fas %.4g e2e %.4g'%(d['value'],d['e2e']['value'],d['cfg1']['value'],d['cfg1']['e2e']['value'],d['fas']['value'],d['fas']['e2e']['value']))
print('ccx resident %.1f ms e2e %.1f ms (nccl %.1f / %.1f) k1 %.1f'%(c['ms_per_step'],c['e2e']['ms_per_step'],c['e2e']['ms_per_step_nccl_gather_every_rank'],c['e2e']['ms_per_step_nccl_gather_rank0_only'],c['roofline']['k1_ms_per_call']))
P
