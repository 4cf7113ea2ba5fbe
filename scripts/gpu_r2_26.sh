#!/bin/bash
# Round 2, GPU call 26 (1 GPU): A/B of small kernel variants on one box (row sums accumulated as sum|J| S, no per-entry L2 prefetch, tracer at 5 blocks per SM).
mkdir -p gpurun_out
for v in base sums noprefetch adv5; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 128 2>&1 | tail -1
done | tee gpurun_out/r2_26_ab.txt
for v in base sums noprefetch adv5 base; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 256 2>&1 | tail -1
done | tee -a gpurun_out/r2_26_ab.txt
