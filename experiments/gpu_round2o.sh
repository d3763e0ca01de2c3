#!/bin/bash
# round 2, call O: code-size sensitivity of the tiled CCX re-scoring kernel (loop unrolling variants), same box
mkdir -p gpurun_out
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
for v in "" _nounroll _unroll6; do
  L=$PWD/detex_b200/_C/libdetex_b200$v.so
  echo "== variant '$v'"
  DETEX_B200_LIB=$L ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ccx_post_tiled" -c 200 --csv --log-file gpurun_out/r2o_ncu$v.csv python /tmp/ccx_only.py > gpurun_out/r2o_log$v.txt 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2o_ncu$v.csv', errors='ignore')))
hdr=None; t=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            if u=='us': v*=1e3
            if u=='ms': v*=1e6
            t.append(v)
print("launches", len(t), "passes3 call: %.1f ms"%(sum(t[:32])/1e6), " passes1 calls: %.1f / %.1f ms"%(sum(t[32:64])/1e6, sum(t[64:96])/1e6))
PY
done
