#!/bin/bash
# Round 2, GPU call 41 (1 GPU): xi_optimal by series / one exp + one reciprocal instead of tanh + divisions: parity of the
# stabilised variants, A/B against the previous build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "su_ or supg" 2>&1 | tail -3 | tee gpurun_out/r2_41_pytest.log
for v in su1 su4 su1 su4; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/prof_su.py 128 2>&1 | tail -1
done | tee gpurun_out/r2_41_ab_su.txt
