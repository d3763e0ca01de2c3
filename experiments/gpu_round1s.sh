#!/bin/bash
# compute-sanitizer on the smoke path (small shapes): memcheck, then racecheck of the smem epilogue
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
