#!/bin/bash
# round 2, call B: every bench section on a reduced station (96 chunks) + the reference arm
mkdir -p gpurun_out
( time python bench.py --steps 1 --warmup 1 --chunks 96 ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "rc=$?" >> gpurun_out/r2b_bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err
tail -20 gpurun_out/r2b_bench.err; head -c 600 gpurun_out/r2b_ref.json
