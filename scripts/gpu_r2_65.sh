#!/bin/bash
# Round 2, GPU call 65 (1 GPU): the stress loop under initcheck (uninitialised device-memory reads).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool initcheck --print-limit 8 python scripts/stress_surface.py 200 > gpurun_out/r2_65_initcheck.log 2>&1; echo "exit $?"
grep -E "ERROR SUMMARY|iterations off|rror:" gpurun_out/r2_65_initcheck.log | tail -5
grep -E "Uninitialized|at .*\+0x|in .*\(|Host Frame.*(cgasm|gather|strip)" gpurun_out/r2_65_initcheck.log | head -40 | cut -c1-220
