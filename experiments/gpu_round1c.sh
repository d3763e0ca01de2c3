#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30
echo "== ccx bench 1024"; timeout 600 python experiments/ccx_bench.py 1024 2>&1 | tail -8
echo "== ccx bench 4096"; timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -5
echo "== bench full"; timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== ncu launch list (our kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dtx -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "== ncu full K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_full -f \
   python bench.py --steps 1 --warmup 1 --chunks 8 --batch 8 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
