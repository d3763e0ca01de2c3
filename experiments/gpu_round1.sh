#!/bin/bash
# One GPU-box session: tests, smoke, bench, ncu launch list, ncu full capture of K1.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log | cut -c1-400
echo "== ncu full K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_full -f \
   python bench.py --steps 1 --warmup 1 --chunks 8 --batch 8 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
