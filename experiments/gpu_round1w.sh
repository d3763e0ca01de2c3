#!/bin/bash
# workflow + decimation on the GPU, K1 chunk-group sweep (DRAM reads vs group size, evict_last signal copies),
# full bench with the final kernels, launch list
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for G in 2 8 16; do
  DTX_K1_GROUP=$G timeout 600 ncu --metrics $M --clock-control none -k regex:k1_kernel -s 1 -c 1 --csv --log-file gpurun_out/k1_traffic_g$G.csv \
     python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/k1_traffic_g$G.log 2>&1
  echo "-- group $G"; grep -E "dram__bytes|gpu__time|lts__" gpurun_out/k1_traffic_g$G.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
echo "== bench full"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
KREG='regex:k0_|k1_kernel|k3_|lta_kernel|basis_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --chunks 96 --no-cpu --no-alt > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
