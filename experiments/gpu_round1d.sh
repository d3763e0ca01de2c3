#!/bin/bash
mkdir -p gpurun_out
KREG='regex:k0_|k1_kernel|k3_kernel|lta_kernel|basis_image|ccx_|stalta|direct_kernel'
echo "== ncu launch list (our kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "== ncu K0/K3 hbm metrics"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k "regex:k0_|k3_kernel|lta_kernel" -s 5 -c 5 --csv --log-file gpurun_out/k0k3.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu > /dev/null 2>&1
wc -l gpurun_out/k0k3.csv
echo "== ncu full K1 at bench batch (48 chunks)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_full48 -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu > gpurun_out/ncu_full48.log 2>&1
tail -2 gpurun_out/ncu_full48.log | cut -c1-200
echo "== ccx launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 2000 --csv --log-file gpurun_out/ccx_launches.csv \
   python experiments/ccx_bench.py 4096 > gpurun_out/ccx_under_ncu.log 2>&1
tail -3 gpurun_out/ccx_under_ncu.log
