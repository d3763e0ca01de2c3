#!/bin/bash
# round 2, call R: after restricting the long first accumulations to long templates: whole suite + smoke + A/B
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -4 gpurun_out/r2r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g e2e %.4g k1 %.1f parity %.3g'%(d['value'], d['e2e']['value'], d['roofline']['k1_ms_per_launch'], d['parity_check']['max_abs_err_vs_fp64']))"
