#!/bin/bash
# Round 2, GPU call 43 (1 GPU): A/B of the carried cross product (one cofactor of the window is the previous step's).
mkdir -p gpurun_out
for c in 128 256; do
  for v in base carry base carry; do
    CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py $c 2>&1 | tail -1
  done
done | tee gpurun_out/r2_43_ab_carry.txt
