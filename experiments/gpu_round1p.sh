#!/bin/bash
# final state of the round: tests, smoke, full bench (default engine + adaptive pass), reference arm,
# ncu launch list, ncu full of K1 (DRAM traffic with 4-chunk groups)
mkdir -p gpurun_out
KREG='regex:k0_|k1_kernel|k3_|lta_kernel|basis_image|ccx_|stalta|direct_kernel|pp_|mag_kernel|corr0'
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench full"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
echo "== ncu launch list (our kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
echo "== ncu full K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_tcgen05_r1p -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/ncu_k1_r1p.log 2>&1
tail -1 gpurun_out/ncu_k1_r1p.log | cut -c1-150
