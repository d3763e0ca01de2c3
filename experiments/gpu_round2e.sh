#!/bin/bash
# round 2, call E: where does the CCX re-scoring time go?  launch list of our kernels + one full capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -q -k "long_array" > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
cat > /tmp/ccx_only.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from detex_b200 import synth
from detex_b200.engine import Engine
X = synth.event_families(3003, 64, 64, 1000, 3, max_shift=100)
eng = Engine(0)
for rep in range(2):
    t0 = time.perf_counter(); r = eng.ccx_condensed(X, 3, engine="tcgen05"); print("ccx s", time.perf_counter() - t0)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dtx|k1_kernel" -c 400 --csv --log-file gpurun_out/r2e_ncu_ccx_launches.csv python /tmp/ccx_only.py > gpurun_out/r2e_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ccx_post_tiled" -s 20 -c 1 -o gpurun_out/r2e_post_tiled python /tmp/ccx_only.py > gpurun_out/r2e_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ccx_scan" -s 20 -c 1 -o gpurun_out/r2e_scan python /tmp/ccx_only.py > gpurun_out/r2e_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2e_ncu1.log
