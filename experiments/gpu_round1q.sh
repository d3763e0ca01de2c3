#!/bin/bash
# other BASELINE configs with the final kernels of round 1 (configs[1], [2], [4]) + HBM metrics of K0/K3
mkdir -p gpurun_out
timeout 900 python experiments/config_runs.py > gpurun_out/config_runs.jsonl 2> gpurun_out/config_runs.err; cat gpurun_out/config_runs.jsonl | cut -c1-400; tail -2 gpurun_out/config_runs.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k "regex:k0_|k3_|lta_kernel" -s 6 -c 6 --csv --log-file gpurun_out/k0k3.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt > /dev/null 2>&1
tail -8 gpurun_out/k0k3.csv | cut -c1-300
