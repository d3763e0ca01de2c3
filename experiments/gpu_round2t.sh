#!/bin/bash
# round 2, call T: ncu evidence for the ring re-scoring (launch list of one configs[2] run, full capture of the ring
# kernel and of the scan)
mkdir -p gpurun_out
CMD="python bench.py --sections ccx --no-cpu --no-alt --chunks 24"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2t_ccx_launches.csv $CMD > gpurun_out/r2t_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ccx_post_ring -s 50 -c 2 -o gpurun_out/r2t_ring $CMD > gpurun_out/r2t_r.log 2>&1
ncu --set full --clock-control none -k regex:ccx_scan -s 50 -c 1 -o gpurun_out/r2t_scan $CMD > gpurun_out/r2t_s.log 2>&1
ls -la gpurun_out/r2t_*
