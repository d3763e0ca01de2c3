#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
for k in 1 2 4; do
echo "== bench kblk=$k chunks=96"; timeout 600 python bench.py --steps 1 --warmup 1 --chunks 96 --kblk $k --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g frac %.4f k1_ms %.1f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['k1_ms_per_launch'], d['clocks']))
"
done
echo "== bench full"; timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1700 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== ncu full K1 (48 chunks)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o gpurun_out/k1_full48 -f \
   python bench.py --steps 1 --warmup 1 --chunks 48 --batch 48 --no-cpu > gpurun_out/ncu_full48.log 2>&1
tail -1 gpurun_out/ncu_full48.log | cut -c1-150
