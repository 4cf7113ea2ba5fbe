#!/bin/bash
# Round 2, GPU call 42 (1 GPU): the occupancy-class test, compute-sanitizer memcheck + racecheck of the class launches and of
# the searched strips on small meshes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "occupancy_classes" 2>&1 | tail -3 | tee gpurun_out/r2_42_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/ab_classes.py 30000 1 > gpurun_out/r2_42_memcheck_classes.log 2>&1; echo "memcheck classes exit $?"; grep -E "ERROR SUMMARY|classes o" gpurun_out/r2_42_memcheck_classes.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/ab_classes.py 30000 1 > gpurun_out/r2_42_racecheck_classes.log 2>&1; echo "racecheck classes exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/r2_42_racecheck_classes.log | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/ab_kernels.py 24 > gpurun_out/r2_42_memcheck_box.log 2>&1; echo "memcheck box exit $?"; grep -E "ERROR SUMMARY" gpurun_out/r2_42_memcheck_box.log | tail -1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/ab_kernels.py 24 > gpurun_out/r2_42_racecheck_box.log 2>&1; echo "racecheck box exit $?"; grep -E "RACECHECK SUMMARY" gpurun_out/r2_42_racecheck_box.log | tail -1
