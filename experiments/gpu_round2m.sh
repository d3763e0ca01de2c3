#!/bin/bash
# round 2, call M: basis image multicast to CTA pairs -- correctness, same-box A/B, L2 -> SM bytes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fused.py tests/test_gpu_ccx.py tests/test_gpu_x8.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -6 gpurun_out/r2m_pytest.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --chunks 96 --no-cpu --no-alt --sections main"
for rep in 1 2; do
  DTX_K1_MC=0 $B > gpurun_out/r2m_mc0_$rep.json 2>> gpurun_out/r2m_err.log
  DTX_K1_MC=1 $B > gpurun_out/r2m_mc1_$rep.json 2>> gpurun_out/r2m_err.log
done
for mc in 0 1; do
DTX_K1_MC=$mc timeout 300 ncu --metrics l1tex__m_xbar2l1tex_read_bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"k1_kernel" -s 1 -c 1 --csv --log-file gpurun_out/r2m_ncu_mc$mc.csv python bench.py --steps 1 --warmup 1 --chunks 48 --no-cpu --no-alt --sections main > gpurun_out/r2m_ncu_mc$mc.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_mc*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'k1 ms', round(d['roofline']['k1_ms_per_launch'],1), 'parity %.3g'%d['parity_check']['max_abs_err_vs_fp64'], 'cands', d['candidates_per_step'], 'hist', d['hist_total'], 'clk', d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -h "k1_kernel" gpurun_out/r2m_ncu_mc*.csv | cut -d, -f5,13- | head; tail -5 gpurun_out/r2m_err.log
