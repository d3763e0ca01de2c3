#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40
for k in 1 2 4; do
echo "== bench kblk=$k chunks=96 batch=48"; timeout 600 python bench.py --steps 1 --warmup 1 --chunks 96 --batch 48 --kblk $k --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g frac %.4f k1share %.3f clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['k1_share_of_step'], d['clocks']))
    else: print(l.rstrip()[:300])
"
done
