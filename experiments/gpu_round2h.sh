#!/bin/bash
# round 2, call H: whole GPU suite + every bench section (reduced station) after the CCX / overlap changes
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
( time python bench.py --steps 2 --warmup 1 --chunks 96 ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "rc=$?" >> gpurun_out/r2h_bench.err
tail -6 gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_smoke.log; tail -6 gpurun_out/r2h_bench.err
