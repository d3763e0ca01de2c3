#!/bin/bash
# CCX configs[2] timing with the final kernels: wall time x3, then the kernel breakdown under ncu
mkdir -p gpurun_out
timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -3
timeout 600 python experiments/ccx_bench.py 4096 2>&1 | tail -2
KREG='regex:k0_|k1_kernel|basis_image|ccx_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 800 --csv --log-file gpurun_out/ccx_launches.csv \
   python experiments/ccx_bench.py 4096 > gpurun_out/ccx_under_ncu.log 2>&1
wc -l gpurun_out/ccx_launches.csv
