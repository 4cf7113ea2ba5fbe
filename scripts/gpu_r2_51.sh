#!/bin/bash
# Round 2, GPU call 51 (1 GPU): 2-D kernels at 6 (momentum, 78 registers) / 8 (tracer, 64 registers) blocks per SM.
mkdir -p gpurun_out
for v in base 2d68 base 2d68; do
  CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py 2048 2d 2>&1 | tail -1
done | tee gpurun_out/r2_51_ab_2d.txt
