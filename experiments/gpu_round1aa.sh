#!/bin/bash
# final-state check: tests, smoke, createCluster at the reference's published workload shape, reference arm
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== createCluster, published shape"; timeout 600 python experiments/createcluster_published_shape.py 2>&1 | tail -3 | tee gpurun_out/createcluster_published.json
