#!/bin/bash
# round 2, call A: full GPU test-suite (incl. BASELINE-scale parity tests) + smoke + a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo skip-bench
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_smoke.log
