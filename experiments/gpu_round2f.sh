#!/bin/bash
# round 2, call F: one-MMA screening series for CCX + leaner tiled re-scoring: correctness, time, launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ccx.py tests/test_gpu_scale.py tests/test_gpu_case1_workflow.py tests/test_workflow.py -m gpu -q -k "ccx or case1 or workflow" > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
cat > /tmp/ccx_only.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from detex_b200 import synth
from detex_b200.engine import Engine
X = synth.event_families(3003, 64, 64, 1000, 3, max_shift=100)
eng = Engine(0)
for passes in (3, 1, 1):
    eng.set_ccx_passes(passes)
    t0 = time.perf_counter(); r = eng.ccx_condensed(X, 3, engine="tcgen05"); dt = time.perf_counter() - t0
    print("passes", passes, "ccx s %.4f" % dt, "k1 ms %.2f" % eng.k1_ms_history().sum())
PY
python /tmp/ccx_only.py > gpurun_out/r2f_ccx_times.log 2>&1; cat gpurun_out/r2f_ccx_times.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ccx|k1_kernel|k0_|basis" -c 2000 --csv --log-file gpurun_out/r2f_ncu_ccx_launches.csv python /tmp/ccx_only.py > gpurun_out/r2f_ncu1.log 2>&1
tail -2 gpurun_out/r2f_ncu1.log
