#!/bin/bash
# round 2, call C: tiled CCX re-scoring (correctness + time) and the A-operand collector reuse A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py -m gpu -q -k "ccx" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
B="python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt"
COLL=$PWD/detex_b200/_C/libdetex_b200_coll.so
DTX_CCX_POST_UNTILED=1 $B --sections ccx > gpurun_out/r2c_ccx_untiled.json 2> gpurun_out/r2c_err.log
$B --sections ccx > gpurun_out/r2c_ccx_tiled.json 2>> gpurun_out/r2c_err.log
DETEX_B200_LIB=$COLL $B --sections ccx > gpurun_out/r2c_ccx_tiled_coll.json 2>> gpurun_out/r2c_err.log
for rep in 1 2; do
  $B --sections main > gpurun_out/r2c_main_default_$rep.json 2>> gpurun_out/r2c_err.log
  DETEX_B200_LIB=$COLL $B --sections main > gpurun_out/r2c_main_coll_$rep.json 2>> gpurun_out/r2c_err.log
done
tail -3 gpurun_out/r2c_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'ccx' in d:
            c=d['ccx']; print(f, 'ccx ms', round(c['ms_per_step'],1), 'e2e ms', round(c['e2e']['ms_per_step'],1), 'k1 ms', round(c['roofline']['k1_ms_per_call'],1), 'frac', round(c['roofline']['frac'],3))
        else:
            print(f, 'value %.4g'%d['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'parity', d['parity_check']['max_abs_err_vs_fp64'], 'clk', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2c_err.log
