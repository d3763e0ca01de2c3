#!/bin/bash
# round 2, call L: drain-warp work A/B on one box (FADD2, pipelined drain), fused mode with the quad path
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_detect.py tests/test_gpu_x8.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log
B="python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
NOF=$PWD/detex_b200/_C/libdetex_b200_nofadd2.so
PIPE=$PWD/detex_b200/_C/libdetex_b200_pipe.so
for rep in 1 2; do
  DETEX_B200_LIB=$NOF $B > gpurun_out/r2l_nofadd2_$rep.json 2>> gpurun_out/r2l_err.log
  $B > gpurun_out/r2l_fadd2_$rep.json 2>> gpurun_out/r2l_err.log
  DETEX_B200_LIB=$PIPE $B > gpurun_out/r2l_pipe_$rep.json 2>> gpurun_out/r2l_err.log
done
$B --kblk 4 > gpurun_out/r2l_fadd2_kblk4.json 2>> gpurun_out/r2l_err.log
DETEX_B200_LIB=$PIPE $B --kblk 4 > gpurun_out/r2l_pipe_kblk4.json 2>> gpurun_out/r2l_err.log
$B --fused > gpurun_out/r2l_fused.json 2>> gpurun_out/r2l_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'share %.4f'%d['roofline']['k1_share_of_step'], 'parity %.3g'%d['parity_check']['max_abs_err_vs_fp64'], 'clk', d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2l_err.log
