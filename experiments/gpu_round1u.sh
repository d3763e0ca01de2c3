#!/bin/bash
# basis statistics + CCX templates on the device: GPU tests, CCX wall time (3 fresh processes), smoke
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2 3; do timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -1; done
timeout 600 python experiments/ccx_bench.py 1024 2>&1 | tail -3
