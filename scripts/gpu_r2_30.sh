#!/bin/bash
# Round 2, GPU call 30 (1 GPU): A/B of the accumulation style (folded into the column registers vs separate dot products) in 2-D, 3-D and the tracer with absorption.
mkdir -p gpurun_out
for v in folded mixed separate folded mixed separate; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 2048 2d 2>&1 | tail -1
done | tee gpurun_out/r2_30_ab.txt
for v in folded mixed separate folded mixed separate; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 128 abs 2>&1 | tail -1
done | tee -a gpurun_out/r2_30_ab.txt
for v in folded separate folded separate; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 256 s3 2>&1 | tail -1
done | tee -a gpurun_out/r2_30_ab.txt
