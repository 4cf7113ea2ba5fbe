#!/bin/bash
# Round 2, GPU call 13 (1 GPU): A/B of kernel variants (reciprocal off the critical path, 3 blocks per SM) on one box.
mkdir -p gpurun_out
for v in base splitsk minb3 splitsk_minb3; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 128 2>&1 | tail -1
done | tee gpurun_out/r2_13_ab.txt
for v in base splitsk; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 256 2>&1 | tail -1
done | tee -a gpurun_out/r2_13_ab.txt
