#!/bin/bash
# model-chosen K1 item order: tests (incl. order invariance), traffic of the default, CCX timing, full bench
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k1_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_traffic_model.csv \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/k1_traffic_model.log 2>&1
echo "-- model-chosen order"; grep -E "dram__bytes|gpu__time" gpurun_out/k1_traffic_model.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
for i in 1 2; do timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -1; done
echo "== bench full"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
