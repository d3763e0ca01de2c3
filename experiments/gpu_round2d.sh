#!/bin/bash
# round 2, call D: scan + tiled CCX re-scoring (correctness + time), long-array test, collector default
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py tests/test_gpu_case1_workflow.py -m gpu -q -k "ccx or long_array or case1" > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
B="python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt"
DTX_CCX_POST_UNTILED=1 $B --sections ccx > gpurun_out/r2d_ccx_untiled.json 2> gpurun_out/r2d_err.log
$B --sections ccx > gpurun_out/r2d_ccx_tiled.json 2>> gpurun_out/r2d_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_ncu_ccx_launches.csv python bench.py --steps 1 --warmup 1 --chunks 24 --no-cpu --no-alt --sections ccx > gpurun_out/r2d_ncu_ccx.log 2>&1
tail -3 gpurun_out/r2d_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_ccx*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['ccx']; print(f, 'ccx ms', round(c['ms_per_step'],1), 'e2e ms', round(c['e2e']['ms_per_step'],1), 'k1 ms', round(c['roofline']['k1_ms_per_call'],1), 'frac', round(c['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2d_err.log
