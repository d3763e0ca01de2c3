#!/bin/bash
# runtime histogram bin count (numBins of _initFAS): full GPU suite, smoke, K3 launch time
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k3_fast -c 6 --csv --log-file gpurun_out/k3_launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > /dev/null 2>&1
grep k3_fast gpurun_out/k3_launches.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160
