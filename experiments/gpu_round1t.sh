#!/bin/bash
# K3 quad fast path: GPU tests, racecheck of the smoke path, launch list for the K3 time, FAS config
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
KREG='regex:k0_|k1_kernel|k3_|lta_kernel|basis_image'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt > gpurun_out/bench_under_ncu.log 2>&1
grep -c k3_fast gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_fast_kernel -s 1 -c 1 -o gpurun_out/k3_full -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/ncu_k3.log 2>&1
tail -1 gpurun_out/ncu_k3.log | cut -c1-120
timeout 600 python experiments/config_runs.py --ccx-events 256 2>/dev/null | tail -1 | cut -c1-400
