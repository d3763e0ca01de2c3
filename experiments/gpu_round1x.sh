#!/bin/bash
# K1 item order: superblocks of basis blocks per wave of tiles (DTX_K1_SUPER) x chunk-group size (DTX_K1_GROUP):
# DRAM reads per 48-chunk launch, then same-box step timing of the candidates
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for cfg in "4 4" "4 8" "8 4" "8 8" "16 8" "48 8" "48 4"; do
  set -- $cfg
  DTX_K1_GROUP=$1 DTX_K1_SUPER=$2 timeout 600 ncu --metrics $M --clock-control none -k regex:k1_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_traffic_g$1_s$2.csv \
     python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/k1_traffic_g$1_s$2.log 2>&1
  echo "-- group $1 super $2"; grep -E "dram__bytes|gpu__time|lts__" gpurun_out/k1_traffic_g$1_s$2.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  grep -o '"max_abs_err_vs_fp64": [0-9.e-]*' gpurun_out/k1_traffic_g$1_s$2.log | head -1
done
for rep in 1 2; do for cfg in "0 0" "8 8" "48 8"; do
  set -- $cfg
  if [ $1 = 0 ]; then unset DTX_K1_GROUP DTX_K1_SUPER; else export DTX_K1_GROUP=$1 DTX_K1_SUPER=$2; fi
  timeout 600 python bench.py --steps 2 --warmup 1 --chunks 192 --no-cpu --no-alt 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('rep $rep group $1 super $2 value %.4g k1_ms %.1f sm_mhz %s' % (d['value'], d['roofline']['k1_ms_per_launch'], d['clocks']['sm_mhz']))
" | tee -a gpurun_out/ab_super.txt
done; done
