#!/bin/bash
# round 2, call N: the full-size bench line (1 GPU, every section) + the reference arm, for profiles/
mkdir -p gpurun_out
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
echo "rc=$?" >> gpurun_out/r2n_bench_n1.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err
tail -5 gpurun_out/r2n_bench_n1.err; tail -4 gpurun_out/r2n_ref.err
