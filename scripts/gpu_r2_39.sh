#!/bin/bash
# Round 2, GPU call 39 (1 GPU): stabilised (SU / SUPG) element kernels with J . inverse(diff) hoisted out of the quadrature loop,
# one division per inverse and per xi: parity tests of the stabilised variants + A/B against the previous build and a
# 3-blocks-per-SM build (168 registers, spills).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "su_ or supg" 2>&1 | tail -3 | tee gpurun_out/r2_39_pytest.log
for v in search su1 su3 search su1 su3; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/prof_su.py 128 2>&1 | tail -1
done | tee gpurun_out/r2_39_ab_su.txt
