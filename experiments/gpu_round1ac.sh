#!/bin/bash
# final state of round 1: full GPU suite, smoke, reference arm, ncu --set full of K1 (model-chosen item order)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 500 gpurun_out/bench_ref.json
echo "== ncu full K1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_final -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu --no-alt > gpurun_out/ncu_k1_final.log 2>&1
ncu -i gpurun_out/k1_final.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; val=rows[2] if len(rows)>2 else rows[1]
want=['gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','launch__registers_per_thread','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor.sum']
for h,u,v in zip(hdr,rows[1],val):
    if h in want: print(h,u,v)
"
