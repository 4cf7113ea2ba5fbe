#!/bin/bash
# Round 2, GPU call 28 (1 GPU): fewer FP64 instructions per pair (row constants of the moments, entries accumulated straight into
# the column registers, tracer right-hand side products at eviction): full suite + A/B against the previous build.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_28_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_28_pytest.log
tail -6 gpurun_out/r2_28_pytest.log
for c in 128 256; do
  for v in base fewer base fewer; do
    CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py $c 2>&1 | tail -1
  done
done | tee gpurun_out/r2_28_ab.txt
