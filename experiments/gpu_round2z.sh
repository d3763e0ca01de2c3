#!/bin/bash
# round 2, call Z: final state of the round on one GPU: whole GPU suite, smoke, full-size bench line (all sections),
# reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -4 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z_bench_n1.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g frac %.3f share %.4f clocks %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['k1_share_of_step'],d['clocks']))
print('cfg1 %.4g fas %.4g ccx %.1f ms e2e %.1f ms frac %.3f'%(d['cfg1']['value'],d['fas']['value'],d['ccx']['ms_per_step'],d['ccx']['e2e']['ms_per_step'],d['ccx']['roofline']['frac']))
print('cpu', d['cpu_baseline']['value'], d['ccx'].get('cpu_baseline',{}).get('pairs_per_s'))
P
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err
tail -1 gpurun_out/r2z_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main > gpurun_out/r2z_ncu.log 2>&1
