#!/bin/bash
# Round 2, GPU call 31 (1 GPU): 2-D momentum with the per-element moment products restored (A/B).
mkdir -p gpurun_out
for v in mixed mixed2 mixed mixed2; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 2048 2d 2>&1 | tail -1
done | tee gpurun_out/r2_31_ab.txt
