#!/bin/bash
# fused 3-lag re-scoring in ccx_post_kernel: CCX tests, timing of the 4096-event matrix, published-shape createCluster
mkdir -p gpurun_out
echo "== pytest ccx + workflow + smoke"; timeout 900 python -m pytest tests/test_gpu_ccx.py tests/test_workflow.py tests/test_gpu_case1_workflow.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -1; done
KREG='regex:k1_kernel|ccx_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/ccx_launches.csv python experiments/ccx_bench.py 4096 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.DictReader(l for l in open('gpurun_out/ccx_launches.csv') if l.startswith('"')))
t=collections.OrderedDict()
for r in rows:
    k=r['Kernel Name'].split('(')[0]
    t.setdefault(k,[0,0.0]); t[k][0]+=1; t[k][1]+=float(r['Metric Value'])/1e6
for k,v in t.items(): print('%-50s %3d launches %10.3f ms' % (k, v[0], v[1]))
PY
timeout 600 python experiments/createcluster_published_shape.py 2>&1 | tail -1 | cut -c1-200
