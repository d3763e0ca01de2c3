#!/bin/bash
# round 2, call G: ncu capture of the tiled CCX re-scoring kernel (single-candidate series: passes = 3)
mkdir -p gpurun_out
cat > /tmp/ccx_only.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from detex_b200 import synth
from detex_b200.engine import Engine
X = synth.event_families(3003, 64, 64, 1000, 3, max_shift=100)
eng = Engine(0)
eng.set_ccx_passes(int(sys.argv[1]))
for rep in range(2):
    t0 = time.perf_counter(); r = eng.ccx_condensed(X, 3, engine="tcgen05"); print("ccx s", time.perf_counter() - t0)
PY
ncu --set full --clock-control none --import-source on -k regex:"ccx_post_tiled" -s 50 -c 1 -o gpurun_out/r2g_post_tiled_p3 python /tmp/ccx_only.py 3 > gpurun_out/r2g_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ccx_post_tiled" -s 50 -c 1 -o gpurun_out/r2g_post_tiled_p1 python /tmp/ccx_only.py 1 > gpurun_out/r2g_ncu2.log 2>&1
ls -la gpurun_out/r2g*.ncu-rep
