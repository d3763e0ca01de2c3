#!/bin/bash
# STA window tests + where do K1's 17 GB of DRAM reads come from?  Variants of the SAME source
# (build flags), one 48-chunk launch each, DRAM / L2 counters only (a few replays, not --set full).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sector_hit_rate.pct
for v in default sighint bothhint nods; do
  case $v in
    default) unset DETEX_B200_LIB;;
    *) export DETEX_B200_LIB=$PWD/detex_b200/_C/libdetex_b200_$v.so;;
  esac
  timeout 600 ncu --metrics $M --clock-control none -k regex:k1_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_traffic_$v.csv \
     python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/k1_traffic_$v.log 2>&1
  echo "-- $v"; grep -E "dram__bytes|gpu__time|lts__" gpurun_out/k1_traffic_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
unset DETEX_B200_LIB
