#!/bin/bash
# Round 2, GPU call 61 (1 GPU): the stress loop under compute-sanitizer memcheck.
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --print-limit 6 python scripts/stress_surface.py 250 > gpurun_out/r2_61_memcheck.log 2>&1; echo "exit $?"
grep -E "ERROR SUMMARY|iterations off|^iteration" gpurun_out/r2_61_memcheck.log | tail -8
grep -E "Invalid|at .*\+0x|by thread|Address|in .*kernel|Host Frame.*cgasm" gpurun_out/r2_61_memcheck.log | head -40
